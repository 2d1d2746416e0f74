#!/bin/bash
# round-2 evidence run on ONE GPU: tests, the driver's bench lines, small/sweep workloads, ncu launch list + full capture of the
# gather kernel, whole-agent numbers.  Everything lands in gpurun_out/r2prof/ and is summarised under profiles/.
O=gpurun_out/r2prof; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
run() { name=$1; shift; timeout 400 python bench.py "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?" >> $O/$name.err; }
run bench_n1
run reference_n1 --impl reference
run bench_c1 --workload c1
run bench_c2 --workload c2
run bench_c5 --workload c5 --out $O/sweep_c5.md
run bench_c4_n1 --workload c4 --no-e2e --no-cpu-baseline
run bench_nature_tc --network nature-tc --no-e2e --no-cpu-baseline
timeout 600 python scripts/full_agent_bench.py > $O/full_agent.md 2> $O/full_agent.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_bulk_kernel -s 3 -c 1 -f -o $O/gather_full python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full.log 2>&1
ncu -i $O/gather_full.ncu-rep --page raw --csv > $O/gather_full_raw.csv 2>/dev/null
ncu -i $O/gather_full.ncu-rep --page details > $O/gather_full_details.txt 2>/dev/null
tail -3 $O/pytest.log; tail -2 $O/smoke.log; for f in $O/*.err; do echo $f; tail -n 1 $f; done
