set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt
timeout 300 python tools/gpu_check.py > gpurun_out/r02a_check.txt 2>&1; echo "check rc $?"
tail -15 gpurun_out/r02a_check.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.txt 2>&1; echo "pytest rc $?"
tail -5 gpurun_out/r02a_pytest.txt
for t in bfgs_slots=0 bfgs_slots=1,bfgs_warps=8 bfgs_slots=1,bfgs_warps=6 bfgs_slots=1,bfgs_warps=4 bfgs_slots=1,bfgs_warps=7 bfgs_slots=1,bfgs_warps=5; do
  timeout 300 python tools/gpu_perf.py --iters 3 --init spectral --tune $t >> gpurun_out/r02a_perf.txt 2>&1; echo "perf $t rc $?"
done
cat gpurun_out/r02a_perf.txt | grep -E "^==|^it"
