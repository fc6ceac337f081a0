cd $GRAFT_REPO_ROOT
for t in bfgs_slots=1,bfgs_warps=8 bfgs_slots=1,bfgs_warps=4; do
  timeout 300 python tools/gpu_perf.py --iters 2 --init spectral --tune $t --lib strutopy_b200/variants/libstm_slt.so --slots-timing 2>&1 | grep -E "^==|^it|rounds"
done
