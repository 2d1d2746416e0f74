#!/usr/bin/env python
"""Narrow down why the prepared pipeline's gather launch is slower than the same launch alone (scripts/gather_probe3.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xagents_b200.hotpath import PPOHotPath  # noqa: E402

dev = torch.device('cuda', 0)
T, E, A = 128, 256, 6


def make(**kw):
    hp = PPOHotPath(T, E, (84, 84, 4), A, device=dev, **kw)
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    hp.obs.copy_(torch.randint(0, 256, hp.obs.shape, dtype=torch.uint8, device=dev, generator=g))
    for n in ('rewards', 'values', 'last_values', 'log_probs', 'actor_out', 'critic_out'):
        getattr(hp, n).normal_()
    hp.dones.zero_()
    hp.actions.copy_(torch.randint(0, A, hp.actions.shape, device=dev).float())
    for k in range(hp.K):
        hp.perms[k].copy_(torch.randperm(hp.N, device=dev).to(torch.int32))
    hp.prepare()
    return hp


def timed(hp, what, reps=12, only_gather=False, losses_after=False):
    evs = []

    def on_gather(i, fn, args):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(hp.data_stream)
        rc = fn(*args)
        b.record(hp.data_stream)
        evs.append((a, b))
        return rc
    for r in range(reps + 3):
        if r == 3:
            evs.clear()
        if only_gather:
            for g in range(hp.n_groups):
                fn, args = hp._gathers[g]
                on_gather(g, fn, args)
            if losses_after:
                hp.compute_stream.wait_stream(hp.data_stream)
                for fn, args in hp._losses:
                    fn(*args)
                hp.data_stream.wait_stream(hp.compute_stream)
        else:
            hp.run(on_gather=on_gather)
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    rows = sum(hp.group_rows) / len(hp.group_rows)
    med = ms[len(ms) // 2]
    print(f'{what}: groups {hp.group_sizes} median {med * 1e3:.1f} us/launch = {(2 * 28224 + 4) * rows / med / 1e6:.0f} GB/s (min {ms[0] * 1e3:.1f} max {ms[-1] * 1e3:.1f})',
          flush=True)


hp = make(sync='event', gather_chunk=16)
timed(hp, 'gather launches only, data stream', only_gather=True)
timed(hp, 'gather launch then the 16 losses', only_gather=True, losses_after=True)
timed(hp, 'hp.run() event sync')
hp2 = make(sync='event', gather_chunk=16, overlap=False)
timed(hp2, 'hp.run() single stream')
hp3 = make(sync='progress', gather_chunk=16)
timed(hp3, 'hp.run() progress sync')
hp4 = make(sync='event')
timed(hp4, 'hp.run() event sync default schedule')
timed(hp4, 'gather launches only default schedule', only_gather=True)
