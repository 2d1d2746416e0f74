#!/bin/bash
O=gpurun_out/s17; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_plan.py tests/test_gpu_round2.py tests/test_gpu_conv.py tests/test_gpu_agents.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 400 python bench.py --network nature-tc --no-e2e --no-cpu-baseline > $O/bench_nature_tc.json 2> $O/bench_nature_tc.err; echo "rc=$?" >> $O/bench_nature_tc.err
timeout 400 python bench.py --network nature-tc --obs-gather on --no-e2e --no-cpu-baseline > $O/bench_nature_tc_gather.json 2> $O/bench_nature_tc_gather.err; echo "rc=$?" >> $O/bench_nature_tc_gather.err
timeout 300 python scripts/cnn_bench.py > $O/cnn_bench.md 2>&1
TRAIN_STEPS=5 timeout 600 python scripts/full_agent_bench.py > $O/full_agent.md 2> $O/full_agent.err
tail -8 $O/pytest.log; tail -3 $O/bench_nature_tc.err; python - <<'PY'
import json
for f in ('bench_nature_tc','bench_nature_tc_gather'):
    try:
        d=json.loads(open(f'gpurun_out/s17/{f}.json').read().strip().splitlines()[-1])
        print(f,{k:d[k] for k in ('value','ms_per_step','ms_per_step_median','host_issue_ms_per_step')})
    except Exception as e: print(f,'ERR',e)
PY
tail -4 $O/cnn_bench.md; head -4 $O/full_agent.md
