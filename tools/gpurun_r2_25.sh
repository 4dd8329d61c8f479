cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/gpu_profile_vae.py 2>&1 | grep -v Warn | tee gpurun_out/r25_vae_profile.log
