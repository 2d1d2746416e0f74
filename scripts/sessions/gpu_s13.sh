#!/bin/bash
O=gpurun_out/s13; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_agents.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
TRAIN_STEPS=5 timeout 600 python scripts/full_agent_bench.py > $O/full_agent.md 2> $O/full_agent.err
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/rollout_launches.csv python scripts/rollout_launches.py > $O/ncu_rollout.log 2>&1
tail -15 $O/pytest.log; head -4 $O/full_agent.md; tail -3 $O/full_agent.err; tail -3 $O/ncu_rollout.log
