#!/usr/bin/env python
"""Fused gradient all-reduce + clip + Adam over peer memory (xagents_b200/peer.py, csrc/peer_adam.cu) under torchrun:
checks it against the oracle's TF-semantics clip_by_global_norm + Adam on the rank-averaged gradient, checks that the
weights are bit-identical on every rank, and times it against the NCCL path (all-reduce + xa_grad_sumsq + xa_clip_adam)
on the Nature CNN's 1.69 M parameters.  Rank 0 prints one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/peer_adam_check.py
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (checker only)
from xagents_b200 import dist as xdist  # noqa: E402
from xagents_b200 import ops, peer  # noqa: E402


def main():
    rank, local_rank, world = xdist.init_from_env()
    assert world > 1, 'run under torchrun with at least 2 ranks'
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    comm = xdist.ShardComm(device=dev)
    out = {'world_size': world, 'transport': peer.transport(comm), 'probe_errors': peer.probe_errors()}
    if out['transport'] is None:
        if rank == 0:
            print(json.dumps(out))
        return
    # ---- parity: 3 updates of 100 003 parameters (not a multiple of 4 * world) ---------------------------------
    n, clip = 100_003, 0.5
    rng = np.random.default_rng(7)
    init = rng.standard_normal(n).astype(np.float32)
    fused = peer.FusedAllReduceAdam(comm, n, init=torch.from_numpy(init).to(dev))
    p_ref, m_ref, v_ref = init.copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    worst = 0.0
    for step in range(1, 4):
        grads = [np.random.default_rng(1000 * step + g).standard_normal(n).astype(np.float32) * (3.0 if step == 2 else 0.01)
                 for g in range(world)]
        fused.grad.zero_()
        fused.grad[:n].copy_(torch.from_numpy(grads[rank]))
        fused.step(step, clip)
        total = grads[0].copy()
        for g in grads[1:]:
            total = total + g                                  # rank order, fp32: the kernel's order
        mean = total * np.float32(1.0 / world)
        clipped = oracle.clip_by_global_norm([mean], clip)[0][0]
        p_ref, m_ref, v_ref = oracle.adam_step(p_ref, clipped, m_ref, v_ref, step)
        got = fused.param[:n].cpu().numpy()
        worst = max(worst, float(np.abs(got - p_ref).max() / np.abs(p_ref).max()))
    everyone = [torch.empty_like(fused.param) for _ in range(world)]
    torch.distributed.all_gather(everyone, fused.param.contiguous())
    out['parity_max_rel_err_vs_oracle'] = worst
    out['weights_identical_across_ranks'] = all(torch.equal(everyone[0], e) for e in everyone[1:])
    out['status'] = fused.status()
    out['parity_ok'] = worst <= 1e-5 and out['weights_identical_across_ranks'] and out['status'] == 0

    # ---- collective C2 through peer memory: 4 all-gathers (both generations twice) against torch.distributed ---------------
    gather = peer.PeerAllGather(comm, 16 * 4)
    ok_c2 = True
    for call in range(4):
        local = torch.randn(16, 4, dtype=torch.float64, device=dev, generator=torch.Generator(dev).manual_seed(100 * call + rank))
        got = torch.empty((world, 16, 4), dtype=torch.float64, device=dev)
        gather.gather(local, got)
        want = [torch.empty_like(local) for _ in range(world)]
        torch.distributed.all_gather(want, local)
        ok_c2 = ok_c2 and torch.equal(got, torch.stack(want))
    out['peer_allgather_ok'] = bool(ok_c2) and gather.status() == 0

    # ---- timing at the Nature CNN's size ------------------------------------------------------------------------
    n = 1_687_719
    big = peer.FusedAllReduceAdam(comm, n)
    big.grad.normal_()
    flat_p = torch.zeros(n + (-n) % 4, device=dev)
    flat_g, m, v = torch.randn_like(flat_p), torch.zeros_like(flat_p), torch.zeros_like(flat_p)
    ws = ops.optim_workspace(dev)

    def nccl_path(step):
        comm.all_reduce_gradients_async(flat_g)
        comm.wait_gradients()
        ops.clip_adam(flat_p, flat_g, m, v, step, workspace=ws, clip_norm=0.5, grad_scale=1.0 / world)

    def timeit(fn, reps=100):
        for s in range(1, 6):
            fn(s)
        comm.barrier()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for s in range(6, 6 + reps):
            fn(s)
        b.record()
        torch.cuda.synchronize(dev)
        return comm.max_over_ranks(a.elapsed_time(b) / reps * 1e3)

    out['fused_us_per_update'] = timeit(lambda s: big.step(s, 0.5))
    out['nccl_allreduce_plus_clip_adam_us_per_update'] = timeit(nccl_path)
    out['status_after_timing'] = big.status()
    out['n_params'] = n
    if rank == 0:
        print(json.dumps(out))
    torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
