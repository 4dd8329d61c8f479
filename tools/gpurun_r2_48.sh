cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_models_gpu.py -q -x 2>&1 | tail -3
GILLB200_BENCH_HF=0 GILLB200_BENCH_C5=0 timeout -k 10 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r48_bench.json 2> gpurun_out/r48_bench.err
tail -c 300 gpurun_out/r48_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r48_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'], d['clocks'], d['ms_per_step']); print(d['stages_ms']); print({k:(v['ms_per_eval']) for k,v in d['unet_eval_breakdown'].items()})
PY
