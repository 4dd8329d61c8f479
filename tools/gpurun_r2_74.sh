cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for t in "convgn gemm2_kernel 2" "convup2 gemm2_kernel 2"; do
  set -- $t
  timeout 300 bash tools/ncu_extract.sh $1 $2 $3 2>&1 | tail -1
done
