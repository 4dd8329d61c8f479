cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 200 python -m pytest tests/test_kernels_gpu.py -x -q -k "groupnorm_fused_into" 2>&1 | tail -12
