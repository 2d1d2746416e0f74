#!/usr/bin/env python
"""Config C5 (BASELINE.json): GAE / n-step / gather microbench sweep, n_steps 5-2048 x n_envs 16-65536.

Prints a markdown report (commit it under profiles/).  Times are CUDA-event averages over back-to-back
launches of prepared ctypes calls (GPU time; the Python wrappers in ops.py add ~15 us of host time per call); shapes whose traffic is < 32 MB stay L2-resident and are latency-bound ("us" is the figure that
matters there, GB/s is flagged).  The CPU column is the oracle's NumPy loop (= the reference's own loop,
xagents/ppo/agent.py:84-93) on the host.
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes  # noqa: E402

import oracle  # noqa: E402
from xagents_b200 import _ffi, ops  # noqa: E402

lib = _ffi.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr())

dev = 'cuda:0'
PEAK = 6542.7


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3          # us


print('# C5 sweep (round 1, 1xB200): returns/GAE scan and permute-gather\n')
print('## GAE (`xa_gae_f32`), algorithmic bytes 16*T*E + 4*E; peak = measured copy 6542.7 GB/s\n')
print('| T | E | MB | auto us | auto GB/s | frac | sequential us | chunked us | n-step auto us | CPU oracle ms | speed-up | note |')
print('|---|---|---|---|---|---|---|---|---|---|---|---|')
for T in (5, 32, 128, 512, 2048):
    for E in (16, 256, 4096, 65536):
        rng = np.random.default_rng(T + E)
        r = rng.standard_normal((T, E)).astype(np.float32)
        v = rng.standard_normal((T, E)).astype(np.float32)
        lv = rng.standard_normal(E).astype(np.float32)
        d = (rng.random((T + 1, E)) < 0.01).astype(np.float32)
        rd, vd, lvd, dd = (torch.as_tensor(x).to(dev) for x in (r, v, lv, d))
        out = torch.empty((T, E), device=dev)
        nbytes = 16 * T * E + 4 * E
        reps = 200 if nbytes < 64e6 else 20
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)   # raw prepared calls: GPU time, not wrapper time
        t = {m: timeit(lambda: lib.xa_gae_f32(P(rd), P(vd), P(lvd), P(dd), P(out), None, T, E, 0.99, 0.95, k, st), reps)
             for k, m in enumerate(('auto', 'sequential', 'chunked'))}
        tn = timeit(lambda: lib.xa_nstep_returns_f32(P(rd), P(dd), P(lvd), P(out), T, E, 0.99, 0, st), reps)
        t0 = time.perf_counter()
        n_cpu = 1 if T * E > 4e6 else 3
        for _ in range(n_cpu):
            oracle.gae_returns(r, d, v, lv, 0.99, 0.95)
        cpu_ms = (time.perf_counter() - t0) / n_cpu * 1e3
        gbs = nbytes / t['auto'] / 1e3
        note = 'L2-resident, latency-bound' if nbytes < 32e6 else ''
        print(f'| {T} | {E} | {nbytes/1e6:.2f} | {t["auto"]:.1f} | {gbs:.0f} | {gbs/PEAK:.2f} | {t["sequential"]:.1f} | {t["chunked"]:.1f} | {tn:.1f} | {cpu_ms:.2f} | {cpu_ms*1e3/t["auto"]:.0f}x | {note} |')

print('\n## Gather (`xa_gather_rows`, 84x84x4 uint8 rows = 28224 B), algorithmic bytes (2F+4) per row\n')
print('| rows | MB moved | bulk us | bulk GB/s | frac | vector us | vector GB/s | torch index_select us | note |')
print('|---|---|---|---|---|---|---|---|---|')
for T, E in ((5, 16), (128, 16), (128, 64), (128, 256), (128, 1024), (128, 4096)):
    N = T * E
    obs = torch.randint(0, 256, (T, E, 84, 84, 4), dtype=torch.uint8, device=dev)
    dst = torch.empty((N, 84, 84, 4), dtype=torch.uint8, device=dev)
    perm = torch.randperm(N, device=dev).to(torch.int32)
    nbytes = (2 * 28224 + 4) * N
    reps = 50 if nbytes < 1e9 else 8
    tb = timeit(lambda: ops.gather_rows(obs, perm, time_major=(T, E), mode='bulk', out=dst), reps)
    tv = timeit(lambda: ops.gather_rows(obs, perm, time_major=(T, E), mode='vector', out=dst), reps)
    flat = obs.view(N, -1)
    rows = ((perm.long() % T) * E + perm.long() // T)
    tt = timeit(lambda: torch.index_select(flat, 0, rows, out=dst.view(N, -1)), reps)
    note = 'fits L2, latency-bound' if nbytes < 100e6 else ''
    print(f'| {N} | {nbytes/1e6:.1f} | {tb:.1f} | {nbytes/tb/1e3:.0f} | {nbytes/tb/1e3/PEAK:.2f} | {tv:.1f} | {nbytes/tv/1e3:.0f} | {tt:.1f} | {note} |')
    del obs, dst
