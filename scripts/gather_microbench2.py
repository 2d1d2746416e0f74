#!/usr/bin/env python
"""Finer sweep of bytes-in-flight per SM for the bulk gather (stages x chunk), 20 reps each."""
import itertools
import os
os.environ.setdefault('XA_TUNING', '1')    # the XA_GATHER_* knobs are honoured only in tuning mode
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xagents_b200 import ops  # noqa: E402

T, E = 128, int(os.environ.get('E', 256))
N = T * E
dev = 'cuda:0'
obs = torch.randint(0, 256, (T, E, 84, 84, 4), dtype=torch.uint8, device=dev)
dst = torch.empty((N, 84, 84, 4), dtype=torch.uint8, device=dev)
perm = torch.randperm(N, device=dev).to(torch.int32)
BYTES = 2 * N * 28224 + 4 * N


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


ms = timeit(lambda: dst.view(-1).copy_(obs.view(-1)))
print(f'torch copy_                               {ms*1e3:8.1f} us  {2*N*28224/ms/1e6:8.1f} GB/s')
ms = timeit(lambda: ops.gather_rows(obs, perm, time_major=(T, E), mode='vector', out=dst))
print(f'vector                                    {ms*1e3:8.1f} us  {BYTES/ms/1e6:8.1f} GB/s')
for (lh, sh) in ((1, 1), (0, 0), (1, 0)):
    for chunk, stages, ctas in itertools.product((28224, 14112, 9408, 7056, 4704), (2, 3, 4, 5, 6, 8), (1, 2)):
        if chunk * stages * ctas > 226 * 1024:
            continue
        kb = chunk * stages * ctas / 1024
        if kb < 20 or kb > 120:
            continue
        os.environ.update(XA_GATHER_STAGES=str(stages), XA_GATHER_CHUNK=str(chunk), XA_GATHER_CTAS=str(ctas),
                          XA_GATHER_LOAD_HINT=str(lh), XA_GATHER_STORE_HINT=str(sh))
        ms = timeit(lambda: ops.gather_rows(obs, perm, time_major=(T, E), mode='bulk', out=dst))
        print(f'bulk hints={lh}{sh} chunk={chunk:6d} stages={stages} ctas={ctas} inflight/SM={kb:6.1f}KB  {ms*1e3:8.1f} us  {BYTES/ms/1e6:8.1f} GB/s')
