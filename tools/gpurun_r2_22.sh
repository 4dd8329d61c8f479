cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 bash tools/ncu_extract.sh attn48 attn4_kernel 2 2>&1 | tail -3
