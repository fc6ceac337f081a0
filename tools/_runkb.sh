cd $GRAFT_REPO_ROOT
for v in ab1 default hu2; do
  L=""; [ $v != default ] && L="--lib strutopy_b200/variants/libstm_$v.so"
  timeout 300 python tools/gpu_perf.py --iters 3 --init spectral $L 2>&1 | grep -E "^==|^it[12]"
done
timeout 300 python tools/gpu_perf.py --iters 2 --init spectral --K 100 --V 20000 --docs 50000 2>&1 | grep -E "^==|^it[1]"
timeout 300 python tools/gpu_perf.py --iters 2 --init spectral --K 100 --V 20000 --docs 50000 --lib strutopy_b200/variants/libstm_ab1.so 2>&1 | grep -E "^==|^it[1]"
timeout 600 python tools/gpu_parity_sweep.py 2>&1 | tail -12 | cut -c1-150
