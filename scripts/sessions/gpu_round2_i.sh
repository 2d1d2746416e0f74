#!/bin/bash
# N GPUs (first arg): 2-GPU tests when N == 2, bench with the fused kernel and with NCCL
N=${1:-2}
O=gpurun_out/r2i_n$N; mkdir -p $O
export NCCL_DEBUG=WARN
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" == "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_agents.py -m gpu -q -k "two or device or shards" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
fi
run() { name=$1; port=$2; shift; shift; timeout 500 $T --master-port $port bench.py --gpus $N "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?" >> $O/$name.err; }
run fused 29522
run nccl 29521 --c1 nccl --no-e2e --no-single-gpu-compare --no-parity-check
run none 29523 --c1 none --no-e2e --no-single-gpu-compare --no-parity-check
timeout 400 $T --master-port 29524 bench.py --gpus $N --impl reference --steps 5 --warmup 1 > $O/reference.json 2> $O/reference.err
[ -f $O/pytest.log ] && tail -3 $O/pytest.log; for f in $O/*.err; do echo $f; tail -n 2 $f; done
python - $O <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1]+'/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],4), d.get('roofline',{}).get('frac'), d.get('roofline',{}).get('whole_step',{}).get('frac_of_nominal_8TBs'), d.get('parity_checked'), d.get('efficiency_same_per_gpu_E'), d.get('efficiency_vs_whole_workload_on_one_gpu'), d.get('single_gpu'))
        if d.get('e2e'): print('   e2e', d['e2e']['value'], d['e2e'].get('h2d_roofline'), d['e2e'].get('frac_of_h2d_roofline'))
    except Exception as e: print(f, 'ERR', e)
PY
