cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/gpu_profile_unet.py 2>&1 | grep -v Warn | tee gpurun_out/r17_unet_profile.log
