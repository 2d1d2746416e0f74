#!/bin/bash
# 1 GPU: progress-sync pipeline: tests + bench variants
O=gpurun_out/r2e; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
run() { name=$1; shift; timeout 300 python bench.py "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?" >> $O/$name.err; }
run n1_default --no-cpu-baseline
run n1_noopt --no-cpu-baseline --no-e2e --no-optimizer
run n1_g88 --no-cpu-baseline --no-e2e --gather-chunk 8
run n1_g4444 --no-cpu-baseline --no-e2e --gather-chunk 4
run n1_c4 --no-cpu-baseline --no-e2e --workload c4
run c1 --workload c1 --no-cpu-baseline
tail -3 $O/pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2e/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],4), d.get('roofline',{}).get('frac'), d['roofline'].get('whole_step',{}).get('frac'), d.get('gae_gather_loss_only',{}).get('ms_per_step'), d['config'].get('minibatches_per_gather_launch'))
    except Exception as e: print(f, 'ERR', e)
PY
