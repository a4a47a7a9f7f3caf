#!/bin/bash
for rep in 1 2; do
for lib in prev new; do
  if [ $lib = prev ]; then export DEWI_B200_LIB=/root/repo/libdewi_prev.so; else unset DEWI_B200_LIB; fi
  python bench.py --rows 12500000 --batch 64 --steps 20 --warmup 3 --sweep=1,8,32 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$lib', 'B64', round(d['roofline']['kernel_ms'],3), round(d['roofline']['frac'],3), 'step', round(d['ms_per_step'],3), '|', [(b['batch'], round(b['kernel_ms'],3), round(b['frac'],3)) for b in d['batch_sweep']])"
done; done
