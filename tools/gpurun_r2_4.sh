cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
GILLB200_GN_APPLY=1 python tools/gpu_norm_bench.py old 2>&1 | grep -v Warn | tee gpurun_out/r4_norm_old.log
GILLB200_GN_APPLY=2 python tools/gpu_norm_bench.py flat 2>&1 | grep -v Warn | tee gpurun_out/r4_norm_flat.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "groupnorm or norm" 2>&1 | tail -3
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r4_bench.json 2> gpurun_out/r4_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r4_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value']); print(d['stages_ms']); print({k:(v['ms_per_eval']) for k,v in d['unet_eval_breakdown'].items()})
PY
