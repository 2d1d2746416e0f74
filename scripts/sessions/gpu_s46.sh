#!/bin/bash
O=gpurun_out/s46; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_plan.py tests/test_gpu_conv.py tests/test_gpu_agents.py tests/test_gpu_parity.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -2 $O/pytest.log
for i in 1 2 3; do timeout 300 python scripts/update_launches.py 2>&1 | tail -1; done > $O/update_eager.log; cat $O/update_eager.log
timeout 300 python scripts/cnn_bench.py 2>&1 | grep "native plan"
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/update_launches.csv python scripts/update_launches.py > $O/ncu_update.log 2>&1
