cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "split_k or wide_pair or conv3x3" 2>&1 | tail -5
timeout 200 python tools/gpu_conv_bench.py r11 2>&1 | grep -v Warn | grep -E "8x8|16x16|totals" | tee gpurun_out/r11_conv.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
GILLB200_BENCH_HF=0 GILLB200_BENCH_C5=0 timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r11_bench.json 2> gpurun_out/r11_bench.err
tail -c 300 gpurun_out/r11_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r11_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value']); print(d['stages_ms']); print({k:(v['ms_per_eval']) for k,v in d['unet_eval_breakdown'].items()})
PY
