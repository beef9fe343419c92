# debug library only (the product library reads no environment): python speaker-recognition-x-vectors_b200/build.py --debug
export XVEC_LIB=$PWD/speaker-recognition-x-vectors_b200/libxvec_b200_debug.so
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -x -q 2>&1 | tail -4
for v in 1 0 1 0; do XVEC_FC_SMALL=$v python bench.py --no-cpu-baseline --no-c5 --no-second-dtype 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fc_small=$v', round(d['value']), 'utt/s', round(d['ms_per_step']*1e3,1), 'us/step  e2e', round(d['e2e']['value']), 'stack', round(d['roofline']['burst']['ms_per_launch']*1e3,1))"; done
