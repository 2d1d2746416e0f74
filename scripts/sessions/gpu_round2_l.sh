#!/bin/bash
O=gpurun_out/r2l; mkdir -p $O
run() { name=$1; shift; timeout 300 python bench.py --workload c4 --n-envs 2048 --no-cpu-baseline --no-e2e --no-optimizer --steps 24 "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?" >> $O/$name.err; }
run prog_97
run prog_88 --gather-chunk 8
run prog_97_static --static-gather
run event_97 --sync event --gather-schedule 9,7
run prog_16x1 --gather-chunk 1
run prog_97_latefork --late-fork
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2l/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        n=len(d['config']['minibatches_per_gather_launch'])
        g=d['gather_launch_ms_each']
        print(f, round(d['ms_per_step'],3), d['config'].get('minibatches_per_gather_launch'), [round(x,1) for x in d['ms_per_step_each']])
        print('    gathers of steps 6..11:', [g[i*n:(i+1)*n] for i in range(6,12)] if n<=4 else 'n/a')
    except Exception as e: print(f, 'ERR', e)
PY
