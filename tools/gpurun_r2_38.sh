cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -q -x -k "conv or vae" 2>&1 | tail -3
timeout -k 10 300 python tools/gpu_profile_vae.py 2>&1 | grep -v Warn | tee gpurun_out/r38_vae_profile.log | head -24
timeout -k 10 300 python tools/gpu_profile_unet.py 2>&1 | grep -v Warn | tee gpurun_out/r38_unet_profile.log | head -12
for t in "convwide gemm2_kernel 2" "conv gemm2_kernel 2"; do
  set -- $t
  timeout 400 bash tools/ncu_extract.sh $1 $2 $3 2>&1 | tail -1
done
