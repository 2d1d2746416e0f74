#!/bin/bash
O=gpurun_out/n2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "two or shard or peer or cuda1 or device" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?" >> $O/bench_n2.err
tail -4 $O/pytest.log; tail -2 $O/bench_n2.err; head -c 700 $O/bench_n2.json
