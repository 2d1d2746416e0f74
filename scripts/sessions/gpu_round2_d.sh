#!/bin/bash
# 2 GPUs: peer-memory probe + fused all-reduce/Adam check, 2-GPU tests, bench N=2 with NCCL and with the fused kernel
O=gpurun_out/r2d; mkdir -p $O
export NCCL_DEBUG=WARN
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $T --master-port 29511 scripts/peer_adam_check.py > $O/peer_check.json 2> $O/peer_check.err; echo "rc=$?" >> $O/peer_check.err
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_agents.py -m gpu -q -k "two or device or shards" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
run() { name=$1; port=$2; shift; shift; timeout 400 $T --master-port $port bench.py --gpus 2 "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?" >> $O/$name.err; }
run n2_nccl 29521 --c1 nccl --no-e2e
run n2_fused 29522 --c1 fused
run n2_none 29523 --c1 none --no-e2e --no-single-gpu-compare --no-parity-check
cat $O/peer_check.json; tail -3 $O/pytest.log; for f in $O/*.err; do echo $f; tail -n 3 $f; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2d/n2*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],4), d.get('roofline',{}).get('frac'), d.get('parity_checked'), d.get('efficiency_same_per_gpu_E'), d.get('efficiency_vs_whole_workload_on_one_gpu'), d.get('single_gpu'))
        if 'e2e' in d: print('   e2e', d['e2e']['value'], d['e2e'].get('h2d_roofline'))
    except Exception as e: print(f, 'ERR', e)
PY
