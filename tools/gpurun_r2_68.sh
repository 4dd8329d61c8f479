cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 200 python -m pytest tests/test_kernels_gpu.py -x -q -k "groupnorm_fused_into or halo or up2" 2>&1 | tail -2
timeout -k 10 200 python tools/gpu_gn_conv_bench.py lookahead 2>&1 | grep -v Warn | cut -c1-170 | tee gpurun_out/r68_gnconv.log
