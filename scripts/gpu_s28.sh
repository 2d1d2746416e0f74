#!/bin/bash
# chained launches: A/B of the policy levels (XA_PDL = 0 never, 1 small + elementwise, 3 small tensor-core / rollout launches only)
O=gpurun_out/s28; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_agents.py tests/test_gpu_plan.py tests/test_gpu_round2.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
for pdl in 1 0 3 1 0 3; do
  export XA_PDL=$pdl
  timeout 300 python bench.py --no-cpu-baseline --no-e2e >> $O/bench_pdl$pdl.json 2>> $O/bench_pdl$pdl.err
  for i in 1 2; do timeout 300 python scripts/update_launches.py 2>&1 | tail -1; done >> $O/update_eager_pdl$pdl.log
done
for pdl in 1 0; do
  export XA_PDL=$pdl
  TRAIN_STEPS=5 timeout 600 python scripts/full_agent_bench.py > $O/full_agent_pdl$pdl.md 2> $O/full_agent_pdl$pdl.err
  timeout 300 python scripts/cnn_bench.py > $O/cnn_bench_pdl$pdl.md 2>&1
done
tail -3 $O/pytest.log
for pdl in 1 0 3; do echo "== XA_PDL=$pdl"; python - <<PY
import json
for line in open('$O/bench_pdl$pdl.json').read().strip().splitlines():
    try:
        d = json.loads(line); print(d['value'], d['ms_per_step'], d.get('roofline', {}).get('frac'))
    except Exception as e: print('ERR', e)
PY
cat $O/update_eager_pdl$pdl.log; done
for pdl in 1 0; do echo "== XA_PDL=$pdl"; head -4 $O/full_agent_pdl$pdl.md; grep "^| 256\|native plan" $O/cnn_bench_pdl$pdl.md; done
