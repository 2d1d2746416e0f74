#!/bin/bash
O=gpurun_out/s52; mkdir -p $O
timeout 400 python bench.py --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?"
python -c "
import json
d = json.loads(open('$O/bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['ms_per_step_median'], d['ms_per_step_max'], d['roofline']['frac'], d['e2e']['value'])"
for i in 1 2; do TRAIN_STEPS=5 timeout 600 python scripts/full_agent_bench.py 2>/dev/null | head -4 | tail -2; done
