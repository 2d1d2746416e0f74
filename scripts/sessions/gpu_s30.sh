#!/bin/bash
# per-kernel DRAM traffic of the benchmarked step with a WARM L2 (--cache-control none): does the optimiser chain's state stay in L2 between minibatches?
O=gpurun_out/s30; mkdir -p $O
timeout 600 ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file $O/launches_warm.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu.log 2>&1
tail -2 $O/ncu.log
