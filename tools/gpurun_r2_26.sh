cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_retrieval_gpu.py -x -q 2>&1 | tail -3
timeout 200 python tools/gpu_topk_bench.py regmerge 2>&1 | grep -v Warn | tee gpurun_out/r26_topk.log
