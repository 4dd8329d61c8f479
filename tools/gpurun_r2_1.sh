set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r1_pytest.log
cat gpurun_out/r1_pytest.log | tail -8
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
tail -c 1500 gpurun_out/r1_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r1_bench.json').read().strip().splitlines()[-1])
    print('value',d['value'],'e2e',d['e2e']['value'],'hf',d.get('hf_eager_gpu'))
    print('retrieval',json.dumps(d['retrieval'])[:1500])
    print('stages',d['stages_ms'])
except Exception as e: print('bench parse failed',e)
PY
