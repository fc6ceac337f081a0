cd $GRAFT_REPO_ROOT
for r in 3 7 1 5; do
  RANK=$r timeout 300 python bench.py --config c5 --steps 10 --warmup 3 --no-cpu > /dev/null 2> gpurun_out/r02_c5_rank$r.err; echo "rank-$r corpus rc $?"; grep -E "StmError" gpurun_out/r02_c5_rank$r.err | head -2
done
