cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 200 python -m pytest tests/test_kernels_gpu.py -x -q -k "groupnorm_fused_into" 2>&1 | tail -2
timeout -k 10 200 python tools/gpu_gn_conv_bench.py w2_3_10_11 2>&1 | grep -v Warn | cut -c1-170 | tee gpurun_out/r67_gnconv.log
GILLB200_GEMM_DEBUG=5 timeout -k 10 200 python tools/gpu_gn_conv_bench.py w2_3_10_11_affine 2>&1 | grep -v Warn | cut -c1-170 | tee -a gpurun_out/r67_gnconv.log
