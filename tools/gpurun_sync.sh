cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
for sel in "test_gemm_stream_k" "test_conv3x3_wide_pair_tile" "test_fused_attention_head_pitch" "test_gemm_epilogues"; do
  echo "=== synccheck $sel" >> gpurun_out/san_sync_detail.log
  timeout 150 $SAN --tool synccheck --print-limit 3 --error-exitcode 0 python -m pytest tests/test_kernels_gpu.py -x -q -k "$sel" 2>&1 | grep -E "=========" | grep -v "Host Frame" | head -14 | cut -c1-200 >> gpurun_out/san_sync_detail.log
done
cat gpurun_out/san_sync_detail.log
GILLB200_FOLD_LN=1 timeout 400 python bench.py --steps 2 --warmup 3 > gpurun_out/r10_bench_foldln.json 2> gpurun_out/r10_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r10_bench_foldln.json').read().strip().splitlines()[-1])
print('FOLD_LN value',d['value'],'e2e',d['e2e']['value']); print(d['stages_ms']); print({k:(v['ms_per_eval']) for k,v in d['unet_eval_breakdown'].items()})
PY
