cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
GILLB200_GN_APPLY=1 timeout 150 python tools/gpu_norm_bench.py old 2>&1 | grep -v Warn | tee gpurun_out/r5_norm_old.log
GILLB200_GN_APPLY=2 timeout 150 python tools/gpu_norm_bench.py flat 2>&1 | grep -v Warn | tee gpurun_out/r5_norm_flat.log
timeout 400 python -m pytest tests/test_kernels_gpu.py -x -q -k "wide_pair or groupnorm or feature_extractor or head_pitch or conv3x3" 2>&1 | tail -6
timeout 200 python tools/gpu_conv_bench.py r5 2>&1 | grep -v Warn | tee gpurun_out/r5_conv.log
GILLB200_ATTN_ISSUER=1 timeout 120 python tools/gpu_attn_bench.py issuer 2>&1 | grep -v Warn | head -4 | tee gpurun_out/r5_attn_issuer.log
timeout 120 python tools/gpu_attn_bench.py default 2>&1 | grep -v Warn | head -4 | tee gpurun_out/r5_attn_default.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r5_bench.json 2> gpurun_out/r5_bench.err
tail -c 400 gpurun_out/r5_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r5_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value']); print(d['stages_ms']); print({k:(v['ms_per_eval']) for k,v in d['unet_eval_breakdown'].items()})
print('c5', json.dumps(d.get('config5_full_surface'))[:800])
PY
