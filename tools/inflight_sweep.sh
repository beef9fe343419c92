for t in 2 3 4 2; do
  XVEC_BENCH_INFLIGHT=$t python bench.py --no-cpu-baseline --no-c5 --no-second-dtype --steps 60 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('inflight=$t', round(d['value']), 'utt/s', round(d['ms_per_step']*1e3,1), 'us/step  e2e', round(d['e2e']['value']))"
done
