set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" 2>&1 | tail -8
GILLB200_ATTN_PTMEM=1 python tools/gpu_attn_bench.py ptmem 2>&1 | tee gpurun_out/r3_attn_ptmem.log
GILLB200_ATTN_PTMEM=0 python tools/gpu_attn_bench.py smemP 2>&1 | tee gpurun_out/r3_attn_smemp.log
GILLB200_ATTN=2 python tools/gpu_attn_bench.py attn2all 2>&1 | tee gpurun_out/r3_attn_attn2all.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err
tail -c 600 gpurun_out/r3_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value']); print(d['stages_ms']); print(json.dumps(d['unet_eval_breakdown']))
PY
