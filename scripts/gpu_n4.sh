#!/bin/bash
O=gpurun_out/n4; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e > $O/bench_n4.json 2> $O/bench_n4.err; echo "rc=$?" >> $O/bench_n4.err
tail -2 $O/bench_n4.err; head -c 600 $O/bench_n4.json
