cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 200 python -m pytest tests/test_kernels_gpu.py -x -q -k "halo or up2 or wide_pair or groupnorm_fused_into" 2>&1 | tail -2
timeout -k 10 200 python tools/gpu_conv_bench.py ahead3 2>&1 | grep -v Warn | cut -c1-80 | tee gpurun_out/r72_conv.log
