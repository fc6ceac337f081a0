cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc $?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "ours rc $?"
timeout 900 python bench.py --config c5 --steps 10 --warmup 3 > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err; echo "c5 rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc $?"
timeout 600 ncu --set full --clock-control none --import-source on --metrics sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed -k regex:'bfgs_kernel|post_group' -s 2 -c 2 -o gpurun_out/prof_r02 python tools/gpu_perf.py --iters 2 --init spectral > gpurun_out/ncu_r02.log 2>&1; echo "ncu full rc $?"
