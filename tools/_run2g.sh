cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_distributed.py -x -q 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2_try.json 2> gpurun_out/r02_bench_n2_try.err; echo "rc $?"; tail -5 gpurun_out/r02_bench_n2_try.err
