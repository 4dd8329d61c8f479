cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "layernorm" 2>&1 | tail -3
timeout -k 10 200 python tools/gpu_small_level.py persistent 2>&1 | grep -v Warn | grep "layernorm rows" | tee gpurun_out/r45_ln.log
GILLB200_LN_PERSISTENT=0 timeout -k 10 200 python tools/gpu_small_level.py oneshot 2>&1 | grep -v Warn | grep "layernorm rows" | tee -a gpurun_out/r45_ln.log
