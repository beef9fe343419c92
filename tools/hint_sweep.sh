# debug library only (the product library reads no environment)
export XVEC_LIB=$PWD/speaker-recognition-x-vectors_b200/libxvec_b200_debug.so
for h in 122 022 002 000 222 102 120; do echo hint $h; XVEC_L2HINT=$h timeout 100 python tools/stack_bench.py --band 0 --iters 40 2>&1 | grep "^band"; done
