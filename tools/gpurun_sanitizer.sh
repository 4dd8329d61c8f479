# compute-sanitizer over the hand-written kernels: memcheck on a broad subset, racecheck / synccheck on the
# kernels with hand-rolled mbarrier / stream-K / TMEM protocols. Summaries land in gpurun_out/ -> profiles/.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
SEL='test_gemm_block_n_variants or test_gemm_epilogues or test_gemm_stream_k or test_conv3x3_stream_k or test_conv3x3_wide_pair_tile or test_fused_attention_head_pitch or test_fused_attention_growing_max or test_small_kernels or test_layernorm or test_groupnorm_nhwc_with_concat or test_geglu_fast_epilogue_fp16'
for tool in memcheck synccheck racecheck; do
  echo "=== $tool" | tee gpurun_out/san_$tool.log
  timeout 900 $SAN --tool $tool --print-limit 20 --error-exitcode 0 python -m pytest tests/test_kernels_gpu.py -x -q -k "$SEL" 2>&1 | grep -v "^$" | tail -40 >> gpurun_out/san_$tool.log
  tail -12 gpurun_out/san_$tool.log
done
echo "=== memcheck retrieval" | tee gpurun_out/san_memcheck_retrieval.log
timeout 600 $SAN --tool memcheck --print-limit 20 --error-exitcode 0 python -m pytest tests/test_retrieval_gpu.py -x -q -k "reference_fixture or per_query or streaming_small" 2>&1 | tail -25 >> gpurun_out/san_memcheck_retrieval.log
tail -8 gpurun_out/san_memcheck_retrieval.log
