cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 100 python tools/gpu_halo_debug.py 2>&1 | grep -v Warn | tail -8
timeout -k 10 240 python -m pytest tests/test_kernels_gpu.py -q -k "conv" 2>&1 | tail -8
timeout -k 10 200 python tools/gpu_conv_bench.py halo 2>&1 | grep -v Warn | tee gpurun_out/r36_conv.log
