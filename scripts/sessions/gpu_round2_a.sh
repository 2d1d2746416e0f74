#!/bin/bash
# round-2 first GPU pass (1 GPU): tests, default bench, small workloads
mkdir -p gpurun_out/r2a
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
timeout 300 python bench.py > gpurun_out/r2a/bench_n1.json 2> gpurun_out/r2a/bench_n1.err; echo "rc=$?" >> gpurun_out/r2a/bench_n1.err
timeout 200 python bench.py --workload c1 > gpurun_out/r2a/bench_c1.json 2> gpurun_out/r2a/bench_c1.err; echo "rc=$?" >> gpurun_out/r2a/bench_c1.err
timeout 200 python bench.py --workload c2 > gpurun_out/r2a/bench_c2.json 2> gpurun_out/r2a/bench_c2.err; echo "rc=$?" >> gpurun_out/r2a/bench_c2.err
timeout 300 python bench.py --workload c5 --out gpurun_out/r2a/sweep_c5.md > gpurun_out/r2a/bench_c5.json 2> gpurun_out/r2a/bench_c5.err; echo "rc=$?" >> gpurun_out/r2a/bench_c5.err
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2a/smoke.log
tail -3 gpurun_out/r2a/pytest.log; tail -2 gpurun_out/r2a/*.err
