#!/bin/bash
O=gpurun_out/r2j; mkdir -p $O
run() { name=$1; shift; timeout 300 python bench.py --workload c4 --n-envs 2048 --no-cpu-baseline --no-e2e --no-optimizer --steps 30 "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?" >> $O/$name.err; }
run progress
run progress_nosampler --no-clock-sampler
run single_stream --no-overlap
run event_chunk1 --sync event --gather-chunk 1
run event_default --sync event
run progress_chunk4 --gather-chunk 4
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), d['config'].get('minibatches_per_gather_launch'), [round(x,1) for x in d['ms_per_step_each']])
    except Exception as e: print(f, 'ERR', e)
PY
