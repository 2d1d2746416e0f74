#!/bin/bash
O=gpurun_out/r2h; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x -k "hotpath or pipeline or c3 or fed" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline --no-e2e "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?" >> $O/$name.err; }
run dyn
run dyn_latefork --late-fork
run static --static-gather
run static_latefork --static-gather --late-fork
run noopt_dyn --no-optimizer
run noopt_static_latefork --no-optimizer --static-gather --late-fork
run dyn_g88 --gather-chunk 8
tail -3 $O/pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2h/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), round(d['roofline']['whole_step']['frac'],4), d.get('gae_gather_loss_only',{}).get('ms_per_step'), d['config'].get('minibatches_per_gather_launch'))
    except Exception as e: print(f, 'ERR', e)
PY
