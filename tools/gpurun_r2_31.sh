cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "groupnorm" 2>&1 | tail -3
timeout 200 python tools/gpu_small_level.py onewave 2>&1 | grep -v Warn | grep "groupnorm(from" | tee gpurun_out/r31_gn.log
GILLB200_GN_ONEWAVE=0 timeout 200 python tools/gpu_small_level.py r1sizing 2>&1 | grep -v Warn | grep "groupnorm(from" | tee -a gpurun_out/r31_gn.log
