cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for t in "conv8 gemm_kernel 2" "gn gn_ 4" "gemm320 gemm_kernel 2"; do
  set -- $t
  timeout 400 bash tools/ncu_extract.sh $1 $2 $3 2>&1 | tail -3
done
