cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "c3 rc $?"; tail -3 gpurun_out/r02_bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --config c5 --steps 10 --warmup 3 > gpurun_out/r02_bench_c5_n8.json 2> gpurun_out/r02_bench_c5_n8.err; echo "c5 rc $?"; tail -3 gpurun_out/r02_bench_c5_n8.err
