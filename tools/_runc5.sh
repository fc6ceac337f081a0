cd $GRAFT_REPO_ROOT
timeout 1200 python bench.py --config c5 --steps 3 --warmup 2 > gpurun_out/r02_bench_c5_try.json 2> gpurun_out/r02_bench_c5_try.err; echo "rc $?"; tail -12 gpurun_out/r02_bench_c5_try.err
