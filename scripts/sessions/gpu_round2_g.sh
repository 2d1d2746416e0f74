#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline --no-e2e "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?" >> $O/$name.err; }
run noopt_event16 --no-optimizer --sync event --gather-chunk 16
run noopt_prog16 --no-optimizer --sync progress --gather-chunk 16
run noopt_event8431 --no-optimizer --sync event
run noopt_prog8431 --no-optimizer --sync progress --gather-schedule 8,4,3,1
run noopt_prog4444 --no-optimizer --sync progress --gather-chunk 4
run opt_prog4444 --sync progress --gather-chunk 4
run opt_prog8431 --sync progress --gather-schedule 8,4,3,1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2g/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],4), d.get('roofline',{}).get('frac'), d['roofline'].get('whole_step',{}).get('frac'), d['roofline']['avg_launch_ms'], d['config'].get('minibatches_per_gather_launch'))
    except Exception as e: print(f, 'ERR', e)
PY
