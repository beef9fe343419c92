for b in 160 170 190; do echo band $b; timeout 200 python tools/power_probe.py bf16 $b 2>&1 | grep "launches     0\|launches  3000"; done
echo "long form 64 x 6000 (1500 m-tiles)"
timeout 200 python tools/stack_bench.py --batch 64 --frames 6000 --band 0,180,240,300 --iters 12 2>&1 | grep "^band"
echo "ragged-like 128k rows: 437 x 300"
timeout 200 python tools/stack_bench.py --batch 437 --frames 300 --band 0,180,256 --iters 20 2>&1 | grep "^band"
