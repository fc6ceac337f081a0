cd $GRAFT_REPO_ROOT
timeout 300 python tools/gpu_perf.py --iters 2 --init spectral --K 100 --V 20000 --docs 50000 2>&1 | grep -E "^==|^it[1]"
timeout 300 python tools/gpu_perf.py --iters 2 --init spectral --K 128 --V 20000 --docs 20000 2>&1 | grep -E "^==|^it[1]"
timeout 300 python tools/gpu_perf.py --iters 2 --init spectral --K 70 --V 20000 --docs 50000 2>&1 | grep -E "^==|^it[1]"
timeout 300 python tools/gpu_perf.py --iters 2 --init spectral 2>&1 | grep -E "^==|^it[1]"
timeout 600 python tools/gpu_parity_sweep.py 2>&1 | tail -6 | cut -c1-170
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
