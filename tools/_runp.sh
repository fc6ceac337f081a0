cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'post_group' -s 1 -c 1 -o gpurun_out/prof_r02_postK100 python tools/gpu_perf.py --iters 2 --init spectral --K 100 --V 20000 --docs 30000 > gpurun_out/ncu_r02_postK100.log 2>&1; tail -1 gpurun_out/ncu_r02_postK100.log
