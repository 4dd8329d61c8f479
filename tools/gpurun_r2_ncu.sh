cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_models_gpu.py -x -q -k "pil_prompts or log_likelihood" 2>&1 | grep -v Warning | tail -15
for t in "attn48 attn3_kernel 2" "xattn48 attn3_kernel 2" "convwide gemm2_kernel 2" "topk topk_scores_kernel 2" "topk1 topk_stream_kernel 2" "geglu gemm_kernel 2"; do
  set -- $t
  timeout 400 bash tools/ncu_extract.sh $1 $2 $3 2>&1 | tail -3
done
ls -la gpurun_out/ | tail -20
