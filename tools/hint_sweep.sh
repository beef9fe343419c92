for h in 122 022 002 000 222 102 120; do echo hint $h; XVEC_L2HINT=$h timeout 100 python tools/stack_bench.py --band 0 --iters 40 2>&1 | grep "^band"; done
