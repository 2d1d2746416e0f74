#!/bin/bash
# N GPUs: the driver's command line (fused C1 by default) and the NCCL variant
N=${1:-8}
O=gpurun_out/r2n_n$N; mkdir -p $O
export NCCL_DEBUG=WARN
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { name=$1; port=$2; shift; shift; timeout 500 $T --master-port $port bench.py --gpus $N "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?" >> $O/$name.err; }
run fused 29522 --steps 20 --warmup 5
run nccl 29521 --c1 nccl --no-e2e --no-single-gpu-compare --no-parity-check
for f in $O/*.err; do echo $f; tail -n 2 $f; done
python - $O <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1]+'/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],4), d.get('roofline',{}).get('frac'), d.get('roofline',{}).get('whole_step',{}).get('frac_of_nominal_8TBs'), d.get('parity_checked'), d.get('efficiency_same_per_gpu_E'), d.get('efficiency_vs_whole_workload_on_one_gpu'), d.get('single_gpu'))
        print('   steps', [round(x,2) for x in d['ms_per_step_each']])
        if d.get('e2e'): print('   e2e', d['e2e']['value'], d['e2e'].get('h2d_roofline'), d['e2e'].get('frac_of_h2d_roofline'))
    except Exception as e: print(f, 'ERR', e)
PY
