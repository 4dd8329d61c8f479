cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in wide_c64 wide_c128_n320 wide_c128_n640 bn160_n320 bn256_n256 bn256_n320 bn128_n128; do
timeout -k 10 100 python tools/gpu_halo_debug.py $c 2>&1 | grep -v Warn | tail -2
done
timeout -k 10 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/gpu_halo_debug.py wide_c128_n640 2>&1 | grep -v Warn | grep -A 25 "=========" | head -60 > gpurun_out/r35_san.log
head -50 gpurun_out/r35_san.log
