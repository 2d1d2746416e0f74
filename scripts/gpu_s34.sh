#!/bin/bash
# CTA pairs (cta_group::2) in the Dense kernel: correctness first (own timeout: a barrier bug traps through the watchdog), then A/B
O=gpurun_out/s37; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x > $O/pytest_gemm.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gemm.log
tail -15 $O/pytest_gemm.log
if grep -q "pytest rc=0" $O/pytest_gemm.log; then
  timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_plan.py tests/test_gpu_agents.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
  tail -3 $O/pytest.log
  for pair in 1 0; do
    export XA_GEMM_TMA_STORE=$pair
    timeout 300 python scripts/gemm_bench.py > $O/gemm_bench_tma$pair.md 2>&1
    timeout 300 python scripts/cnn_bench.py > $O/cnn_bench_tma$pair.md 2>&1
    for i in 1 2; do timeout 300 python scripts/update_launches.py 2>&1 | tail -1; done > $O/update_eager_tma$pair.log
    echo "== XA_GEMM_TMA_STORE=$pair"; tail -9 $O/gemm_bench_tma$pair.md | head -7; grep "native plan" $O/cnn_bench_tma$pair.md; cat $O/update_eager_tma$pair.log
  done
fi
