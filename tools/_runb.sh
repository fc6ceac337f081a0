cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
M=sm__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__sass_thread_inst_executed_op_dfma_pred_on.sum,sm__sass_thread_inst_executed_op_dmul_pred_on.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.sum,sm__ops_path_tensor_src_fp64.sum,sm__sass_thread_inst_executed_op_fp64_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum
timeout 600 ncu --clock-control none --metrics $M -k regex:'bfgs_kernel|post_group' -s 2 -c 2 python tools/gpu_perf.py --iters 2 --init spectral 2>&1 | grep -E "stm::|sass_thread|dmma|tensor_src|dram__|gpu__time|inst_executed" > gpurun_out/r02_fp64_counts.txt
cat gpurun_out/r02_fp64_counts.txt
