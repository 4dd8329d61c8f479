cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/r13_bench_${N}gpu.json 2> gpurun_out/r13_bench_${N}gpu.err
tail -c 1500 gpurun_out/r13_bench_${N}gpu.err | grep -v Warning
python - $N <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r13_bench_{n}gpu.json').read().strip().splitlines()[-1])
    print('value',d['value'],'e2e',d['e2e']['value'])
    r=d['retrieval']; print('retrieval qps',r['value'],'ms',r['ms_per_batch'],'phases',r['phases_ms'],'graph',r['cuda_graph'],'checks',r['checks'])
    print('c5', json.dumps(d.get('config5_full_surface'))[:500])
except Exception as e: print('bench parse failed',e)
PY
