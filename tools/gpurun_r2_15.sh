cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -5
timeout 200 python tools/gpu_gemm_bench.py r15 2>&1 | grep -v Warn | grep -E "M65536|totals" | tee gpurun_out/r15_gemm.log
GILLB200_BRES=0 timeout 200 python tools/gpu_gemm_bench.py r15nobres 2>&1 | grep -v Warn | grep -E "M65536|totals" | tee gpurun_out/r15_gemm_nobres.log
timeout 120 python tools/gpu_mapper_profile.py 2>&1 | grep -v Warn | tee gpurun_out/r15_mapper.log | head -9
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
GILLB200_BENCH_HF=0 GILLB200_BENCH_C5=0 timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r15_bench.json 2> gpurun_out/r15_bench.err
tail -c 300 gpurun_out/r15_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r15_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value']); print(d['stages_ms']); print({k:(v['ms_per_eval']) for k,v in d['unet_eval_breakdown'].items()})
print('mapper', d['mapper']['ms_per_batch'], d['mapper']['rel_err_vs_fp64_oracle_B4'])
PY
