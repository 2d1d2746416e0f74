#!/bin/bash
# 1 GPU: failed tests again, bench variants after the chain fix, ncu launch list of the default bench
O=gpurun_out/r2c; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_conv.py -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
run() { name=$1; shift; timeout 300 python bench.py "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?" >> $O/$name.err; }
run n1_default --no-cpu-baseline --no-e2e
run n1_noopt --no-cpu-baseline --no-e2e --no-optimizer
run n1_sched_4444 --no-cpu-baseline --no-e2e --gather-chunk 4
run n1_sched_44431_s3 --no-cpu-baseline --no-e2e --staging 3 --gather-schedule 4,4,4,3,1
run n1_sched_6442_s3 --no-cpu-baseline --no-e2e --staging 3 --gather-schedule 6,4,4,1,1
run c2 --workload c2 --no-cpu-baseline
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu.log 2>&1
tail -3 $O/pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],4), d.get('roofline',{}).get('frac'), d.get('gae_gather_loss_only',{}).get('ms_per_step'), d.get('env_major_reorder_80_frames_us'))
    except Exception as e: print(f, 'ERR', e)
PY
