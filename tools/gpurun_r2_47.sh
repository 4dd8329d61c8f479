cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
GILLB200_EPI_STG=2 timeout -k 10 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm or conv or geglu or layernorm_folded or groupnorm_from" 2>&1 | tail -3
GILLB200_EPI_STG=1 timeout -k 10 120 python tools/gpu_geglu_modes.py stg 2>&1 | grep -v Warn | tee gpurun_out/r47_stg.log
GILLB200_EPI_STG=1 timeout -k 10 200 python tools/gpu_gemm_bench.py stg 2>&1 | grep -v Warn | cut -c1-64 | tee -a gpurun_out/r47_stg.log
