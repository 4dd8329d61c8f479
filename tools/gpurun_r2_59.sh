cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 200 python tools/gpu_gn_conv_bench.py full 2>&1 | grep -v Warn | tee gpurun_out/r59_gnconv.log
GILLB200_GEMM_DEBUG=4 timeout -k 10 200 python tools/gpu_gn_conv_bench.py sync_only 2>&1 | grep -v Warn | tee -a gpurun_out/r59_gnconv.log
