#!/bin/bash
O=gpurun_out/s18; mkdir -p $O
timeout 300 python scripts/update_launches.py > $O/update_eager.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/update_launches.csv python scripts/update_launches.py > $O/ncu.log 2>&1
tail -2 $O/update_eager.log; tail -2 $O/ncu.log; wc -l $O/update_launches.csv
