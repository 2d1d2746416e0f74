#!/bin/bash
O=gpurun_out/r2m; mkdir -p $O
run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 24 "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?" >> $O/$name.err; }
run e2048_noopt --workload c4 --n-envs 2048 --no-optimizer
run e2048 --workload c4 --n-envs 2048
run e2048_16x1 --workload c4 --n-envs 2048 --gather-chunk 1
run c4 --workload c4
run c3
run c3_noopt --no-optimizer
XA_TUNING=1 XA_GATHER_SPREAD=0 timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 24 --workload c4 --n-envs 2048 > $O/e2048_nospread.json 2> $O/e2048_nospread.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2m/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        n=len(d['config']['minibatches_per_gather_launch'])
        g=d['gather_launch_ms_each']
        print(f, round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), round(d['roofline']['whole_step']['frac'],4), d['config'].get('minibatches_per_gather_launch'), [round(x,1) for x in d['ms_per_step_each']][:14])
    except Exception as e: print(f, 'ERR', e)
PY
