cd $GRAFT_REPO_ROOT
for c in 3 4 5; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --tune host_chunks=$c 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('chunks $c value %.3fM e2e %.3fM ratio %.3f' % (d['value']/1e6, d['e2e']['value']/1e6, d['e2e']['value']/d['value']))"
done
