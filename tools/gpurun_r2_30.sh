cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r30_launches_full.csv python tools/ncu_launch_list.py > gpurun_out/r30_ncu.log 2>&1
tail -2 gpurun_out/r30_ncu.log
python tools/ncu_summarize_launches.py gpurun_out/r30_launches_full.csv gpurun_out/r30_launches_summary.csv | head -40
