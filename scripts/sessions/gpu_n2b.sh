#!/bin/bash
O=gpurun_out/n2c; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "two or shard or peer or cuda1 or device" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?" >> $O/bench_n2.err
tail -3 $O/pytest.log; tail -2 $O/bench_n2.err
python -c "
import json
d = json.loads(open('$O/bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['efficiency_same_per_gpu_E'], d['efficiency_vs_whole_workload_on_one_gpu'], d['parity_checked'], d['single_gpu'])"
