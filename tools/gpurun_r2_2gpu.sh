set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
tail -c 2000 gpurun_out/r2_bench_2gpu.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_2gpu.json').read().strip().splitlines()[-1])
    print('value',d['value'],'e2e',d['e2e']['value'])
    print('retrieval',json.dumps(d['retrieval'])[:2500])
except Exception as e: print('bench parse failed',e)
PY
