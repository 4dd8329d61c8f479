cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 200 python -m pytest tests/test_kernels_gpu.py -x -q -k "groupnorm_fused_into" 2>&1 | tail -3
timeout -k 10 200 python tools/gpu_gn_conv_bench.py four_warps 2>&1 | grep -v Warn | tee gpurun_out/r60_gnconv.log
