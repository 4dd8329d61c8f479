cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "up2 or halo or wide_pair" 2>&1 | tail -6
timeout -k 10 900 python -m pytest tests/test_models_gpu.py -x -q -k "unet or vae or sd_pipe" 2>&1 | tail -3
timeout -k 10 300 python tools/gpu_profile_unet.py 2>&1 | grep -v Warn | tee gpurun_out/r54_unet_profile.log | grep -E "total|conv3x3|upsample|up2" | head -12
timeout -k 10 300 python tools/gpu_profile_vae.py 2>&1 | grep -v Warn | tee gpurun_out/r54_vae_profile.log | grep -E "eager|conv3x3|upsample|up2" | head -12
