#!/bin/bash
O=gpurun_out/s44; mkdir -p $O
for cap in 1184 592 296 148 1184 296; do
  XA_OPT_BLOCKS=$cap timeout 300 python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('cap $cap', d['value'], d['ms_per_step'], d['ms_per_step_median'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])" >> $O/ab.log
done
cat $O/ab.log
