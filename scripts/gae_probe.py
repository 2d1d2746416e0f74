#!/usr/bin/env python
"""GAE vs n-step scan at a large shape, timed with raw prepared ctypes calls (no wrapper overhead)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xagents_b200 import _ffi  # noqa: E402

T, E = int(os.environ.get('T', 128)), int(os.environ.get('E', 65536))
PAD = int(os.environ.get('PAD', 0))
dev = 'cuda:0'
lib = _ffi.lib()


def alloc(rows):
    t = torch.randn(rows * E + PAD, device=dev)
    return t[PAD:] if PAD else t


r, v, d, out, adv = alloc(T), alloc(T), (torch.rand((T + 1) * E, device=dev) < 0.01).float(), alloc(T), alloc(T)
lv = torch.randn(E, device=dev)
P = lambda t: ctypes.c_void_p(t.data_ptr())
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


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
    return a.elapsed_time(b) / reps * 1e3


for mode, name in ((0, 'auto'), (1, 'sequential'), (2, 'chunked')):
    us = timeit(lambda: lib.xa_gae_f32(P(r), P(v), P(lv), P(d), P(out), None, T, E, 0.99, 0.95, mode, s))
    print(f'gae   {name:10s} T={T} E={E}: {us:8.1f} us  {(16*T*E+4*E)/us/1e3:7.0f} GB/s')
    us = timeit(lambda: lib.xa_gae_f32(P(r), P(v), P(lv), P(d), P(out), P(adv), T, E, 0.99, 0.95, mode, s))
    print(f'gae+A {name:10s} T={T} E={E}: {us:8.1f} us  {(20*T*E+4*E)/us/1e3:7.0f} GB/s')
    us = timeit(lambda: lib.xa_nstep_returns_f32(P(r), P(d), P(lv), P(out), T, E, 0.99, mode, s))
    print(f'nstep {name:10s} T={T} E={E}: {us:8.1f} us  {(12*T*E+4*E)/us/1e3:7.0f} GB/s')
