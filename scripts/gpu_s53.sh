#!/bin/bash
O=gpurun_out/s53; mkdir -p $O
for w in c1 c2; do timeout 200 python bench.py --workload $w --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err; echo "$w rc=$?"; head -c 700 $O/bench_$w.json; echo; tail -2 $O/bench_$w.err; done
