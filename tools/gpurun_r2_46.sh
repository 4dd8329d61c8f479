cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 200 python -m pytest tests/test_kernels_gpu.py -x -q -k "attn or attention" 2>&1 | tail -4
timeout -k 10 120 python tools/gpu_attn_bench.py a4q 2>&1 | grep -v Warn | grep "Lk77" | tee gpurun_out/r46_attn.log
GILLB200_ATTN4Q=0 timeout -k 10 120 python tools/gpu_attn_bench.py a4 2>&1 | grep -v Warn | grep "Lk77" | tee -a gpurun_out/r46_attn.log
