#!/usr/bin/env python
"""How long one gradient all-reduce (1,687,719 fp32) takes on this box, alone and under a running gather."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xagents_b200 import dist as xdist  # noqa: E402
from xagents_b200 import ops  # noqa: E402

rank, local, world = xdist.init_from_env()
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
g = torch.zeros(1_687_719, device=dev)


def timed(fn, reps):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


t_alone = timed(lambda: dist.all_reduce(g), 100)
T, E = 128, 512
obs = torch.randint(0, 256, (T, E, 84, 84, 4), dtype=torch.uint8, device=dev)
dst = torch.empty((T * E, 84, 84, 4), dtype=torch.uint8, device=dev)
perm = torch.randperm(T * E, device=dev).to(torch.int32)
side = torch.cuda.Stream(dev, priority=-1 if os.environ.get('HIPRI') else 0)
t_gather = timed(lambda: ops.gather_rows(obs, perm, time_major=(T, E), out=dst), 5)


def both():
    with torch.cuda.stream(side):
        for _ in range(4):
            dist.all_reduce(g)
    ops.gather_rows(obs, perm, time_major=(T, E), out=dst)
    torch.cuda.current_stream().wait_stream(side)


t_both = timed(both, 5)
if rank == 0:
    print(f'world={world}: all_reduce alone {t_alone:.1f} us; epoch gather alone {t_gather:.1f} us; '
          f'gather + 4 concurrent all_reduces {t_both:.1f} us')
dist.destroy_process_group()
