#!/bin/bash
# chained launches (programmatic dependent launch): full GPU suite, then A/B of the live numbers with XA_PDL=1 / 0
O=gpurun_out/s27; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
for pdl in 1 0; do
  export XA_PDL=$pdl
  timeout 300 python bench.py --no-cpu-baseline --no-e2e > $O/bench_pdl$pdl.json 2> $O/bench_pdl$pdl.err; echo "rc=$?" >> $O/bench_pdl$pdl.err
  timeout 300 python bench.py --network nature-tc --no-e2e --no-cpu-baseline > $O/bench_tc_pdl$pdl.json 2> $O/bench_tc_pdl$pdl.err; echo "rc=$?" >> $O/bench_tc_pdl$pdl.err
  for i in 1 2; do timeout 300 python scripts/update_launches.py 2>&1 | tail -1; done > $O/update_eager_pdl$pdl.log
  TRAIN_STEPS=5 timeout 600 python scripts/full_agent_bench.py > $O/full_agent_pdl$pdl.md 2> $O/full_agent_pdl$pdl.err
  timeout 300 python scripts/cnn_bench.py > $O/cnn_bench_pdl$pdl.md 2>&1
done
tail -3 $O/pytest.log
for pdl in 1 0; do echo "== XA_PDL=$pdl"; python - <<PY
import json
for f in ('$O/bench_pdl$pdl.json', '$O/bench_tc_pdl$pdl.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d.get('roofline', {}).get('frac'))
    except Exception as e: print(f, 'ERR', e)
PY
cat $O/update_eager_pdl$pdl.log; head -8 $O/full_agent_pdl$pdl.md; grep "^| 256\|native plan" $O/cnn_bench_pdl$pdl.md; done
