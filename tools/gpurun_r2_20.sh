cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python tools/gpu_attn_bench.py attn4 2>&1 | grep -v Warn | grep "pitch48" | tee gpurun_out/r20_attn.log
GILLB200_ATTN4=0 timeout 120 python tools/gpu_attn_bench.py attn3 2>&1 | grep -v Warn | grep "pitch48" | tee -a gpurun_out/r20_attn.log
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attn or attention" 2>&1 | tail -5
