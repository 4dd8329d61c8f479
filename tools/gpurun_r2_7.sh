cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_models_gpu.py -x -q -k "pil_prompts or mapper or log_likelihood" 2>&1 | grep -v Warning | tail -30
timeout 120 python tools/gpu_mapper_profile.py 2>&1 | grep -v Warn | tee gpurun_out/r7_mapper.log | head -12
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "geglu or gemm" 2>&1 | tail -4
timeout 200 python tools/gpu_gemm_bench.py r7 2>&1 | grep -v Warn | grep -E "geglu|totals|N1280 K5120|N320 K1280" | tee gpurun_out/r7_gemm.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r7_bench.json 2> gpurun_out/r7_bench.err
tail -c 300 gpurun_out/r7_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r7_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value']); print(d['stages_ms']); print({k:(v['ms_per_eval']) for k,v in d['unet_eval_breakdown'].items()})
print('mapper', json.dumps(d['mapper'])[:600])
PY
