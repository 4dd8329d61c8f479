cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_kernels_gpu.py -q -x 2>&1 | tail -3
timeout -k 10 200 python tools/gpu_gemm_bench.py nbuf4 2>&1 | grep -v Warn | grep "res" | cut -c1-64 | tee gpurun_out/r49_res.log
GILLB200_EPI_RES_NBUF=3 timeout -k 10 200 python tools/gpu_gemm_bench.py nbuf3 2>&1 | grep -v Warn | grep "res" | cut -c1-64 | tee -a gpurun_out/r49_res.log
timeout -k 10 200 python tools/gpu_conv_bench.py nbuf4 2>&1 | grep -v Warn | grep -E "totals|64x64 C320->320|32x32 C640->640" | cut -c1-90 | tee -a gpurun_out/r49_res.log
GILLB200_EPI_RES_NBUF=3 timeout -k 10 200 python tools/gpu_conv_bench.py nbuf3 2>&1 | grep -v Warn | grep -E "totals|64x64 C320->320|32x32 C640->640" | cut -c1-90 | tee -a gpurun_out/r49_res.log
