cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout -k 10 200 python tools/gpu_conv_bench.py halo2 2>&1 | grep -v Warn | grep -E "auto" | cut -c1-80 | tee gpurun_out/r37_conv.log
timeout -k 10 300 python tools/gpu_profile_vae.py 2>&1 | grep -v Warn | head -14
GILLB200_BENCH_HF=0 GILLB200_BENCH_C5=0 timeout -k 10 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r37_bench.json 2> gpurun_out/r37_bench.err
tail -c 300 gpurun_out/r37_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r37_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'], d['clocks']); print(d['stages_ms']); print({k:(v['ms_per_eval']) for k,v in d['unet_eval_breakdown'].items()})
PY
