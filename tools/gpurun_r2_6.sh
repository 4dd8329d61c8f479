cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "wide_pair or conv3x3" 2>&1 | tail -4
timeout 200 python tools/gpu_gemm_bench.py r6 2>&1 | grep -v Warn | tee gpurun_out/r6_gemm.log
timeout 120 python tools/gpu_mapper_profile.py 2>&1 | grep -v Warn | tee gpurun_out/r6_mapper.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r6_bench.json 2> gpurun_out/r6_bench.err
tail -c 300 gpurun_out/r6_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r6_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value']); print(d['stages_ms']); print({k:(v['ms_per_eval']) for k,v in d['unet_eval_breakdown'].items()})
print('roofline', d['roofline'])
PY
