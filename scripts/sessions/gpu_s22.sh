#!/bin/bash
O=gpurun_out/s23; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_plan.py tests/test_gpu_agents.py tests/test_gpu_round2.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 300 python scripts/cnn_bench.py > $O/cnn_bench.md 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/cnn_launches.csv python scripts/cnn_launches.py > $O/ncu_launches.log 2>&1
timeout 300 python scripts/update_launches.py > $O/update_eager.log 2>&1
tail -12 $O/pytest.log; tail -6 $O/cnn_bench.md; tail -1 $O/update_eager.log
