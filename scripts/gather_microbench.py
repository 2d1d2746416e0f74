#!/usr/bin/env python
"""Gather kernel in isolation: back-to-back launches timed with one CUDA-event pair, swept over the
XA_GATHER_* knobs, random vs identity permutation, bulk vs vector path, and a torch copy_ for scale."""
import itertools
import os
os.environ.setdefault('XA_TUNING', '1')    # the XA_GATHER_* knobs are honoured only in tuning mode
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xagents_b200 import ops  # noqa: E402

T, E = 128, 256
N = T * E
dev = 'cuda:0'
obs = torch.randint(0, 256, (T, E, 84, 84, 4), dtype=torch.uint8, device=dev)
dst = torch.empty((N, 84, 84, 4), dtype=torch.uint8, device=dev)
perm = torch.randperm(N, device=dev).to(torch.int32)
ident_flat = torch.arange(N, device=dev, dtype=torch.int32)
BYTES = 2 * N * 28224 + 4 * N


def timeit(fn, reps=10):
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
print(f'torch copy_ 925MB                         {ms*1e3:8.1f} us  {2*N*28224/ms/1e6:8.1f} GB/s')
for mode, idx, tm, label in (('bulk', perm, (T, E), 'random/env-major'), ('bulk', ident_flat, None, 'identity/flat'),
                             ('vector', perm, (T, E), 'random/env-major'), ('vector', ident_flat, None, 'identity/flat')):
    src = obs if tm else obs.view(N, 84, 84, 4)
    ms = timeit(lambda: ops.gather_rows(src, idx, time_major=tm, mode=mode, out=dst))
    print(f'{mode:6s} {label:18s} default knobs      {ms*1e3:8.1f} us  {BYTES/ms/1e6:8.1f} GB/s')
grid = itertools.product((4, 6, 8), (56 * 1024, 14112, 9408), (1, 2, 4), (0, 1), (0, 1))
for stages, chunk, ctas, lh, sh in grid:
    if chunk * stages * ctas > 226 * 1024 or (chunk > 28224 // 2 + 16 and ctas > 1 and stages * 28224 * ctas > 226 * 1024):
        continue
    os.environ.update(XA_GATHER_STAGES=str(stages), XA_GATHER_CHUNK=str(chunk), XA_GATHER_CTAS=str(ctas),
                      XA_GATHER_LOAD_HINT=str(lh), XA_GATHER_STORE_HINT=str(sh))
    ms = timeit(lambda: ops.gather_rows(obs, perm, time_major=(T, E), mode='bulk', out=dst), reps=6)
    print(f'bulk stages={stages} chunk={chunk:6d} ctas={ctas} load_hint={lh} store_hint={sh}   {ms*1e3:8.1f} us  {BYTES/ms/1e6:8.1f} GB/s')
