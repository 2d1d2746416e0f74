#!/bin/bash
O=gpurun_out/s45; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "adam or optim or clip or hotpath" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -2 $O/pytest.log
for i in 1 2 3; do
  timeout 300 python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('keep', d['value'], d['ms_per_step'], d['ms_per_step_median'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])" >> $O/ab.log
done
cat $O/ab.log
