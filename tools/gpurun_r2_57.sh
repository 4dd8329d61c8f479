cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_models_gpu.py -x -q -k "unet or vae or sd_pipe" 2>&1 | tail -3
for v in 1 0; do
GILLB200_FUSED_GN=$v GILLB200_BENCH_HF=0 GILLB200_BENCH_C5=0 timeout -k 10 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r57_bench_$v.json 2> gpurun_out/r57_bench.err
python - $v <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r57_bench_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('FUSED_GN',sys.argv[1],'value',d['value'],'e2e',d['e2e']['value'], d['clocks']['sm_mhz'], d['stages_ms'])
print({k:(v['ms_per_eval']) for k,v in d['unet_eval_breakdown'].items()})
PY
done
