cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "narrow or groupnorm" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_models_gpu.py -x -q -k "unet or vae or sd_pipe" 2>&1 | tail -3
timeout 200 python tools/gpu_small_level.py packedsilu 2>&1 | grep -v Warn | grep "groupnorm(from" | tee gpurun_out/r29_gn.log
timeout 300 python tools/gpu_profile_unet.py 2>&1 | grep -v Warn | tee gpurun_out/r29_unet_profile.log | head -12
timeout 300 python tools/gpu_profile_vae.py 2>&1 | grep -v Warn | tee gpurun_out/r29_vae_profile.log | head -12
