#!/bin/bash
# round-2 second GPU pass (1 GPU): tests + bench variants
O=gpurun_out/r2b; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
run() { name=$1; shift; timeout 300 python bench.py "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?" >> $O/$name.err; }
run n1_default --no-cpu-baseline
run n1_staging3 --no-cpu-baseline --no-e2e --staging 3
run n1_noopt --no-cpu-baseline --no-e2e --no-optimizer
run n1_sched_8431_s3 --no-cpu-baseline --no-e2e --staging 3 --gather-schedule 8,4,3,1
run n1_sched_4444 --no-cpu-baseline --no-e2e --gather-chunk 4
run n1_sched_44431_s3 --no-cpu-baseline --no-e2e --staging 3 --gather-schedule 4,4,4,3,1
run c1 --workload c1 --no-cpu-baseline
run c2 --workload c2 --no-cpu-baseline
tail -3 $O/pytest.log; for f in $O/*.err; do echo $f; tail -n 2 $f; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],4), d.get('roofline',{}).get('frac'), d.get('gae_gather_loss_only',{}).get('ms_per_step'), d.get('env_major_reorder_80_frames_us'))
    except Exception as e: print(f, 'ERR', e)
PY
