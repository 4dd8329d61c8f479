cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attn or attention" 2>&1 | tail -3
: > gpurun_out/r27_attn.log
for p in 3 0 4; do
GILLB200_ATTN_POLY=$p timeout 120 python tools/gpu_attn_bench.py a4w_poly$p 2>&1 | grep -v Warn | grep "pitch96\|pitch48" | tee -a gpurun_out/r27_attn.log
done
GILLB200_ATTN4=1 timeout 120 python tools/gpu_attn_bench.py attn2 2>&1 | grep -v Warn | grep "pitch96" | tee -a gpurun_out/r27_attn.log
