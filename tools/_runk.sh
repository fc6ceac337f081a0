cd $GRAFT_REPO_ROOT
timeout 1200 python bench.py --config c5 --docs 500000 --steps 8 --warmup 3 --no-cpu > gpurun_out/r02_c5_500k.json 2> gpurun_out/r02_c5_500k.err; echo rc $?; grep -E "StmError" gpurun_out/r02_c5_500k.err | head -3
python -c "
import json; d=json.loads(open('gpurun_out/r02_c5_500k.json').read().strip().split(chr(10))[-1]); print(d['value'], d['ms_per_step'], d['breakdown_ms'], d['elbo_trace_tail'])"
timeout 600 python -m pytest tests/test_mnreg.py -m gpu -x -q 2>&1 | tail -3
