#!/usr/bin/env python
"""tcgen05 GEMM (xa_gemm_bf16_tn) vs cuBLAS (torch.matmul, bf16) on the network's dense-layer shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xagents_b200 import ops  # noqa: E402

dev = 'cuda:0'


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


print('| M | N | K | what | ours us | ours TFLOP/s | cuBLAS us | cuBLAS TFLOP/s | ratio |')
print('|---|---|---|---|---|---|---|---|---|')
for m, n, k, what in [(8192, 512, 3136, 'fc fwd (B=8192)'), (8192, 3136, 512, 'fc dgrad'), (512, 3136, 8192, 'fc wgrad'),
                      (65536, 512, 3136, 'fc fwd (B=65536)'), (8192, 8, 512, 'heads fwd'), (8192, 8192, 8192, 'square 8192')]:
    a = torch.randn((m, k), device=dev).to(torch.bfloat16)
    b = torch.randn((n, k), device=dev).to(torch.bfloat16)
    out = torch.empty((m, n), device=dev, dtype=torch.bfloat16)
    t_ours = timeit(lambda: ops.gemm_bf16_tn(a, b, out=out))
    torch.cuda.synchronize()
    bt = b.t()
    t_ref = timeit(lambda: torch.matmul(a, bt, out=out))
    fl = 2.0 * m * n * k
    print(f'| {m} | {n} | {k} | {what} | {t_ours:.1f} | {fl/t_ours/1e6:.0f} | {t_ref:.1f} | {fl/t_ref/1e6:.0f} | {t_ref/t_ours:.2f} |')

# the network's own weight-gradient product: dW = dY^T X from the row-major tensors as the other kernels leave them
# (xa_gemm_bf16_atb, MN-major UMMA operands) against cuBLAS on the transposed view; and the heads as the fused kernel
import ctypes  # noqa: E402
from xagents_b200 import _ffi  # noqa: E402
B = 8192
dy = torch.randn((B, 512), device=dev).to(torch.bfloat16)
xx = torch.randn((B, 3136), device=dev).to(torch.bfloat16)
t_ours = timeit(lambda: ops.gemm_bf16_atb(dy, xx))
outw = torch.empty((512, 3136), device=dev, dtype=torch.bfloat16)
t_ref = timeit(lambda: torch.matmul(dy.t(), xx, out=outw))
fl = 2.0 * B * 512 * 3136
print(f'| 512 | 3136 | {B} | fc wgrad as the network runs it (A^T B, row-major operands; incl. split reduction) | {t_ours:.1f} | {fl/t_ours/1e6:.0f} | '
      f'{t_ref:.1f} | {fl/t_ref/1e6:.0f} | {t_ref/t_ours:.2f} |')
h = torch.randn((B, 512), device=dev).relu().to(torch.bfloat16)
wh = torch.randn((8, 512), device=dev).to(torch.bfloat16)
bh = torch.zeros(8, device=dev)
actor, critic = torch.empty((B, 6), device=dev), torch.empty(B, device=dev)
lib = _ffi.lib()
p_ = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
t_ours = timeit(lambda: lib.xa_heads_forward_bf16(p_(h), p_(wh), p_(bh), p_(actor), p_(critic), B, 512, 6, st))
o8 = torch.empty((B, 8), device=dev, dtype=torch.bfloat16)
t_ref = timeit(lambda: torch.matmul(h, wh.t(), out=o8))
print(f'| {B} | 8 | 512 | heads fwd as the network runs it (fused CUDA-core kernel, fp32 outputs + bias) | {t_ours:.1f} | - | {t_ref:.1f} | - | {t_ref/t_ours:.2f} |')
