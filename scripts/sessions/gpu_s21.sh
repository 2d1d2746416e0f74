#!/bin/bash
O=gpurun_out/s21; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_plan.py tests/test_gpu_round2.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 300 python scripts/cnn_bench.py 2>&1 | tail -3 > $O/cnn_tail.txt
timeout 300 python scripts/update_launches.py > $O/update_eager.log 2>&1
timeout 300 python scripts/update_launches.py gather > $O/update_eager_gather.log 2>&1
timeout 400 python bench.py --network nature-tc --no-e2e --no-cpu-baseline > $O/bench_nature_tc.json 2> $O/bench_nature_tc.err; echo "rc=$?" >> $O/bench_nature_tc.err
TRAIN_STEPS=5 timeout 600 python scripts/full_agent_bench.py > $O/full_agent.md 2> $O/full_agent.err
tail -3 $O/pytest.log; cat $O/cnn_tail.txt; tail -1 $O/update_eager.log; tail -1 $O/update_eager_gather.log; head -4 $O/full_agent.md
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s21/bench_nature_tc.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','ms_per_step_median','host_issue_ms_per_step')})
PY
