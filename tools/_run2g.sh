cd $GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --config c5 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_bench_c5_n2.json 2> gpurun_out/r02_bench_c5_n2.err; echo "c5 rc $?"; grep -E "StmError|converge" gpurun_out/r02_bench_c5_n2.err | head -4
