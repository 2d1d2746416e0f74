#!/bin/bash
O=gpurun_out/n8; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e > $O/bench_n8.json 2> $O/bench_n8.err; echo "rc=$?" >> $O/bench_n8.err
tail -2 $O/bench_n8.err
python -c "
import json
d = json.loads(open('$O/bench_n8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['efficiency_same_per_gpu_E'], d['efficiency_vs_whole_workload_on_one_gpu'], d['parity_checked'], d['single_gpu'])"
