cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attn4 or narrow" 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r32_bench.json 2> gpurun_out/r32_bench.err
tail -c 300 gpurun_out/r32_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r32_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'], d['clocks']); print(d['stages_ms']); print({k:(v['ms_per_eval']) for k,v in d['unet_eval_breakdown'].items()})
print('mapper', d['mapper']['ms_per_batch'], d['mapper']['rel_err_vs_fp64_oracle_B4'], 'retr', d['retrieval']['value'])
print('hf', d['hf_eager_gpu']); print('c5', d['config5_full_surface'])
PY
