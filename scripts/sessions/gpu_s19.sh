#!/bin/bash
O=gpurun_out/s19; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_plan.py tests/test_gpu_round2.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 300 python scripts/cnn_bench.py > $O/cnn_bench.md 2>&1
timeout 300 python scripts/update_launches.py > $O/update_eager.log 2>&1
timeout 300 python scripts/update_launches.py gather > $O/update_eager_gather.log 2>&1
tail -3 $O/pytest.log; tail -5 $O/cnn_bench.md; tail -1 $O/update_eager.log; tail -1 $O/update_eager_gather.log
