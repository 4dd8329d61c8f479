cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r70_launches_full.csv python tools/ncu_launch_list.py > gpurun_out/r70_ncu.log 2>&1
python tools/ncu_summarize_launches.py gpurun_out/r70_launches_full.csv gpurun_out/r70_launches_summary.csv | head -30
timeout -k 10 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r70_bench.json 2> gpurun_out/r70_bench.err
tail -c 300 gpurun_out/r70_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r70_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'], d['clocks'], 'launches', d['gpu_launches']); print(d['stages_ms']); print({k:(v['ms_per_eval']) for k,v in d['unet_eval_breakdown'].items()})
print('roofline', d['roofline']['kernel'], d['roofline']['achieved'], d['roofline']['frac']); print(d['rooflines_by_family'])
print('mapper', d['mapper']['ms_per_batch'], 'retr', d['retrieval']['value'], 'hf', d['hf_eager_gpu']['value'], d['hf_eager_gpu']['ours_over_hf_eager_e2e'], 'c5', d['config5_full_surface']['value'])
PY
