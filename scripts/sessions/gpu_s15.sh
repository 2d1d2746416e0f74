#!/bin/bash
O=gpurun_out/s16; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_plan.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 300 python scripts/cnn_bench.py > $O/cnn_bench.md 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/cnn_launches.csv python scripts/cnn_launches.py > $O/ncu_launches.log 2>&1
tail -5 $O/pytest.log; tail -5 $O/cnn_bench.md; grep -E "heads|finalize" $O/cnn_launches.csv | cut -d, -f5,12-15
