# sustained (power-capped) throughput of the stack kernel for several band heights: tools/power_probe.py <dtype> <band>
for b in ${BANDS:-0 180 200 225 250}; do echo band $b; timeout 200 python tools/power_probe.py bf16 $b 2>&1 | grep "launches     0\|launches  3000"; done
