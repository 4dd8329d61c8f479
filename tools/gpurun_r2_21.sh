cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
: > gpurun_out/r21_attn.log
for p in 0 6 4 3 2; do
GILLB200_ATTN_POLY=$p timeout 120 python tools/gpu_attn_bench.py poly$p 2>&1 | grep -v Warn | grep "pitch48" | tee -a gpurun_out/r21_attn.log
done
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attn or attention" 2>&1 | tail -5
