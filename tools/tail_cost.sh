# What the small tail kernels (pool_finalize, segment6 GEMM + split-K reduce) cost a step: bench `value` with them skipped
# (debug library only: python speaker-recognition-x-vectors_b200/build.py --debug)
export XVEC_LIB=$PWD/speaker-recognition-x-vectors_b200/libxvec_b200_debug.so
for t in 0 1 2 0; do
  XVEC_SKIP_TAIL=$t python bench.py --no-cpu-baseline --no-c5 --no-second-dtype --steps 50 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('skip_tail=$t', round(d['value']), 'utt/s', round(d['ms_per_step']*1e3,1), 'us/step  e2e', round(d['e2e']['value']))"
done
