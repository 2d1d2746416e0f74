#!/bin/bash
# session evidence run: full GPU test suite, bench lines, network tables, launch list of one forward+backward
O=gpurun_out/s12; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
timeout 400 python bench.py --network nature-tc --no-e2e --no-cpu-baseline > $O/bench_nature_tc.json 2> $O/bench_nature_tc.err; echo "rc=$?" >> $O/bench_nature_tc.err
timeout 300 python scripts/gemm_bench.py > $O/gemm_bench.md 2>&1
timeout 300 python scripts/cnn_bench.py > $O/cnn_bench.md 2>&1
timeout 600 python scripts/full_agent_bench.py > $O/full_agent.md 2> $O/full_agent.err
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/cnn_launches.csv python scripts/cnn_launches.py > $O/ncu_launches.log 2>&1
tail -3 $O/pytest.log; tail -2 $O/smoke.log; for f in $O/*.err; do echo $f; tail -n 2 $f; done; cat $O/bench_n1.json | head -c 600; echo; cat $O/bench_nature_tc.json | head -c 900; echo; tail -4 $O/gemm_bench.md; tail -4 $O/cnn_bench.md; head -12 $O/full_agent.md
