#!/bin/bash
# tuning sweep of the CTA-pair kernel's ring depths (one process per setting; small corpus)
for cfg in "0 8" "0 4" "0 3"; do
  set -- $cfg
  DEWI_TC2_STAGES=$2 python bench.py --rows 12500000 --batch 4096 --steps 3 --warmup 2 --sweep=1024 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('q=$1 e=$2', 'B4096', round(d['roofline']['achieved']), 'TF', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'], '| B1024', [round(b['achieved']) for b in d['batch_sweep']])"
done
