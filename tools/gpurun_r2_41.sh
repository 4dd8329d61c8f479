cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_kernels_gpu.py -q -x 2>&1 | tail -3
timeout -k 10 120 python tools/gpu_geglu_modes.py pipe 2>&1 | grep -v Warn | tee gpurun_out/r41_geglu.log
timeout -k 10 200 python tools/gpu_gemm_bench.py pipe 2>&1 | grep -v Warn | cut -c1-70 | tee gpurun_out/r41_gemm.log
