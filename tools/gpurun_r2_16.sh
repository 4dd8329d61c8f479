cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_retrieval_gpu.py -x -q 2>&1 | tail -4
timeout 200 python tools/gpu_topk_bench.py warm 2>&1 | grep -v Warn | tee gpurun_out/r16_topk_warm.log
GILLB200_TOPK_WARM=0 timeout 200 python tools/gpu_topk_bench.py cold 2>&1 | grep -v Warn | tee gpurun_out/r16_topk_cold.log
timeout 300 python -m pytest tests/test_models_gpu.py -x -q -k "mapper" 2>&1 | tail -3
