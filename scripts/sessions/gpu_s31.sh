#!/bin/bash
O=gpurun_out/s31; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 300 python scripts/cnn_bench.py > $O/cnn_bench.md 2>&1
TRAIN_STEPS=5 timeout 600 python scripts/full_agent_bench.py > $O/full_agent.md 2> $O/full_agent.err
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/rollout_launches.csv python scripts/rollout_launches.py > $O/ncu_rollout.log 2>&1
tail -3 $O/pytest.log; head -5 $O/cnn_bench.md; head -4 $O/full_agent.md
