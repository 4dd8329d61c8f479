cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "stream or conv" 2>&1 | tail -3
timeout 200 python tools/gpu_small_level.py normal 2>&1 | grep -v Warn | tee gpurun_out/r19b_small.log
GILLB200_GEMM_DEBUG=1 timeout 200 python tools/gpu_small_level.py mma_only 2>&1 | grep -v Warn | tee -a gpurun_out/r19b_small.log
