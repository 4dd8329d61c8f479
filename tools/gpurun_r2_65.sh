cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 200 python -m pytest tests/test_kernels_gpu.py -x -q -k "geglu or gemm_epilogues" 2>&1 | tail -3
timeout -k 10 120 python tools/gpu_geglu_modes.py estrin 2>&1 | grep -v Warn | grep geglu | tee gpurun_out/r65_geglu.log
