# compute-sanitizer over the hand-written kernels: memcheck on a broad subset, racecheck / synccheck on the
# kernels with hand-rolled mbarrier / stream-K / TMEM protocols. Summaries land in gpurun_out/ -> profiles/.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
SEL='test_gemm_epilogues or test_gemm_stream_k or test_conv3x3_wide_pair_tile or test_conv3x3_halo_tile or test_conv3x3_up2 or test_conv3x3_with_groupnorm_fused or test_conv3x3_narrow or test_layernorm_persistent or test_attn4q or test_fused_attention_head_pitch or test_attn4 or test_small_kernels or test_groupnorm_nhwc_with_concat or test_groupnorm_from_epilogue or test_geglu_fast_epilogue_fp16'
run() { # tool, seconds, test file, -k selection
  echo "=== $1: pytest $3 -k '$4'" > gpurun_out/san_$1_$5.log
  timeout $2 $SAN --tool $1 --print-limit 10 --error-exitcode 0 python -m pytest $3 -x -q -k "$4" 2>&1 | grep -v "^$" | grep -E "ERROR SUMMARY|passed|failed|Error|error|hazard|Hazard|=========" | tail -25 >> gpurun_out/san_$1_$5.log
  echo "(exit/timeout status of the sanitizer pipeline: ${PIPESTATUS[0]})" >> gpurun_out/san_$1_$5.log
  tail -6 gpurun_out/san_$1_$5.log
}
run memcheck 420 tests/test_kernels_gpu.py "$SEL" kernels
run memcheck 200 tests/test_retrieval_gpu.py "reference_fixture or per_query or streaming_small" retrieval
#run synccheck 240 tests/test_kernels_gpu.py "test_gemm_stream_k or test_conv3x3_wide_pair_tile or test_conv3x3_halo_tile or test_fused_attention_head_pitch or test_attn4" kernels
run racecheck 300 tests/test_kernels_gpu.py "test_conv3x3_with_groupnorm_fused or test_conv3x3_up2 or test_attn4q" kernels
