#!/usr/bin/env python
"""Why is a gather launch slower inside bench.py than alone?  Event-timed xa_gather_minibatch launches (bulk path), no other
work on the GPU: rows per launch x source footprint x launch spacing x an NVML sampler thread polling beside it."""
import ctypes
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from xagents_b200 import _ffi  # noqa: E402

lib = _ffi.lib()
dev = torch.device('cuda', 0)
P = lambda t: ctypes.c_void_p(t.data_ptr())
F = 28224


def run(T, E, epochs, reps=10, sampler=False, gap_ms=0.0, progress=False):
    N = T * E
    obs = torch.randint(0, 256, (T, E, 84, 84, 4), dtype=torch.uint8, device=dev)
    perms = torch.stack([torch.randperm(N, device=dev).to(torch.int32) for _ in range(epochs)]).view(-1)
    rows = perms.numel()
    dst = torch.empty((rows, F), dtype=torch.uint8, device=dev)
    prog = torch.zeros(64, dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream(dev)
    st = ctypes.c_void_p(s.cuda_stream)
    smp = None
    if sampler:
        smp = bench.ClockSampler(0)
        smp.start()
        time.sleep(0.2)

    def launch():
        if progress:
            rc = lib.xa_gather_rows_progress(P(obs), P(perms), P(dst), rows, F, N, T, E, P(prog), 0, N, max(1, N // 4), None, st)
        else:
            rc = lib.xa_gather_rows(P(obs), P(perms), P(dst), rows, F, N, T, E, 1, st)
        assert rc == 0, lib.xa_last_error()
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(s)
        launch()
        b.record(s)
        if gap_ms:
            torch.cuda.synchronize()
            time.sleep(gap_ms * 1e-3)
    torch.cuda.synchronize()
    if smp:
        smp.stop()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    med = ms[len(ms) // 2]
    gbs = (2 * F + 4) * rows / med / 1e6
    print(f'T={T} E={E} epochs={epochs} rows/launch={rows} sampler={sampler} gap_ms={gap_ms} progress={progress}: median {med * 1e3:.1f} us  {gbs:.0f} GB/s '
          f'(min {ms[0] * 1e3:.1f}, max {ms[-1] * 1e3:.1f})', flush=True)
    del obs, dst


for kw in (dict(T=128, E=256, epochs=1), dict(T=128, E=256, epochs=2), dict(T=128, E=256, epochs=4), dict(T=128, E=1024, epochs=1),
           dict(T=128, E=256, epochs=4, sampler=True), dict(T=128, E=256, epochs=1, sampler=True),
           dict(T=128, E=256, epochs=4, gap_ms=2.0), dict(T=128, E=256, epochs=4, progress=True), dict(T=128, E=256, epochs=1, progress=True),
           dict(T=128, E=2048, epochs=1), dict(T=128, E=2048, epochs=1, progress=True)):
    run(**kw)
