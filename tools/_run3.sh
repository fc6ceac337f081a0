cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bfgs_slots' -s 1 -c 1 -o gpurun_out/prof_r02a_slots python tools/gpu_perf.py --iters 2 --init spectral --tune bfgs_slots=1,bfgs_warps=8 > gpurun_out/ncu_r02a.log 2>&1
tail -3 gpurun_out/ncu_r02a.log
