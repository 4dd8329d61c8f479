cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for m in 0 1 2; do GILLB200_GEMM_DEBUG=$m timeout -k 10 120 python tools/gpu_geglu_modes.py dbg$m 2>&1 | grep -v Warn | tee -a gpurun_out/r39_geglu_modes.log; done
