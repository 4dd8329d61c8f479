cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r12_launches.csv python tools/gpu_one_conv.py 8 1280 1280 > /dev/null 2>&1
grep -E "gemm|splitk|gn_|elementwise" gpurun_out/r12_launches.csv | awk -F'","' '{print $5, $(NF)}' | tail -12
GILLB200_SPLITK=0 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r12_launches_sk.csv python tools/gpu_one_conv.py 8 1280 1280 > /dev/null 2>&1
grep -E "gemm|splitk" gpurun_out/r12_launches_sk.csv | awk -F'","' '{print $5, $(NF)}' | tail -6
