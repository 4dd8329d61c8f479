cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
GILLB200_BENCH_HF=0 GILLB200_BENCH_C5=0 timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r28_bench.json 2> gpurun_out/r28_bench.err
tail -c 300 gpurun_out/r28_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r28_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'], d['clocks']); print(d['stages_ms']); print({k:(v['ms_per_eval']) for k,v in d['unet_eval_breakdown'].items()})
print('mapper', d['mapper']['ms_per_batch'], d['mapper']['rel_err_vs_fp64_oracle_B4'], 'retr', d['retrieval']['value'])
PY
