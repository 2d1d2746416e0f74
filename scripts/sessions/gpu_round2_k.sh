#!/bin/bash
O=gpurun_out/r2k; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x -k "hotpath or pipeline or c3 or fed" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 30 "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?" >> $O/$name.err; }
run e2048_noopt --workload c4 --n-envs 2048 --no-optimizer
run e2048 --workload c4 --n-envs 2048
run e512 --workload c4 --n-envs 512
run c3
run c4 --workload c4
tail -3 $O/pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2k/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), round(d['roofline']['whole_step']['frac'],4), d['config'].get('minibatches_per_gather_launch'), (d.get('gae_gather_loss_only') or {}).get('ms_per_step'), [round(x,2) for x in d['ms_per_step_each']][:12])
    except Exception as e: print(f, 'ERR', e)
PY
