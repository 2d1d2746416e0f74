#!/bin/bash
O=gpurun_out/s11; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 400 python bench.py --network nature-tc --no-e2e --no-cpu-baseline > $O/bench_nature_tc.json 2> $O/bench_nature_tc.err; echo "rc=$?" >> $O/bench_nature_tc.err
TRAIN_STEPS=5 timeout 600 python scripts/full_agent_bench.py > $O/full_agent.md 2> $O/full_agent.err
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/rollout_launches.csv python scripts/rollout_launches.py > $O/ncu_rollout.log 2>&1
tail -3 $O/pytest.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/s11/bench_nature_tc.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','ms_per_step_median','host_issue_ms_per_step')})
PY
head -4 $O/full_agent.md; tail -3 $O/ncu_rollout.log
