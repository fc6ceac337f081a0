cd $GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "c3 rc $?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --config c5 --steps 10 --warmup 3 > gpurun_out/r02_bench_c5_n8.json 2> gpurun_out/r02_bench_c5_n8.err; echo "c5 rc $?"; grep -E "StmError|Error" gpurun_out/r02_bench_c5_n8.err | head -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; echo "n4 rc $?"
timeout 600 python -m pytest tests/test_gpu_distributed.py -x -q 2>&1 | tail -2
