#!/bin/bash
# 2 GPUs, 512 envs per GPU (the per-minibatch budget of C4 on 8 GPUs): what the per-minibatch chain costs under load
N=2
O=gpurun_out/r2o; mkdir -p $O
export NCCL_DEBUG=WARN
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { name=$1; port=$2; shift; shift; timeout 300 $T --master-port $port bench.py --gpus $N --workload c4 --n-envs 1024 --no-e2e --no-single-gpu-compare --no-parity-check --time-updates "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?" >> $O/$name.err; }
run fused 29522
run nccl 29521 --c1 nccl
run none 29523 --c1 none
run noopt 29524 --no-optimizer
for f in $O/*.err; do echo $f; tail -n 1 $f; done
python - $O <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1]+'/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3), d.get('update_ms'), [round(x,2) for x in d['ms_per_step_each']][:8])
    except Exception as e: print(f, 'ERR', e)
PY
