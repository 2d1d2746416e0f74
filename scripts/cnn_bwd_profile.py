#!/usr/bin/env python
"""Per-operation timing of the tensor-core network's backward pass at B = 8192."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xagents_b200 import ops  # noqa: E402
from xagents_b200.agents import NatureCnnTc  # noqa: E402

dev = 'cuda:0'
B = int(os.environ.get('B', 8192))


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


net = NatureCnnTc(4, 6).cuda().refresh()
op = net._op
x = torch.randint(0, 256, (B, 84, 84, 4), dtype=torch.uint8, device=dev)
x1 = ops.space_to_depth_u8_bf16(x, 4)
x2 = ops.conv2d_nhwc_bf16(x1, op.w1, 2, 2, bias=op.b1, relu=True, out_s2d=True)
x3 = ops.conv2d_nhwc_bf16(x2, op.w2, 2, 2, bias=op.b2, relu=True)
y3 = ops.conv2d_nhwc_bf16(x3, op.w3, 3, 3, bias=op.b3, relu=True)
h = ops.gemm_bf16_tn(y3.view(B, -1), op.wf, bias=op.bf_, relu=True, out_dtype=torch.bfloat16)
dh = torch.randn_like(h)
dy3 = torch.randn_like(y3)
dy2 = torch.randn_like(x3)
dy1 = torch.randn_like(x2)
rows = []
t = lambda name, fn: rows.append((name, timeit(fn)))
t('fc wgrad: transposes', lambda: (ops.transpose_bf16(dh), ops.transpose_bf16(y3.view(B, -1))))
dht, y3t = ops.transpose_bf16(dh), ops.transpose_bf16(y3.view(B, -1))
t('fc wgrad: gemm', lambda: ops.gemm_bf16_tn(dht, y3t))
t('fc dgrad: gemm+mask', lambda: ops.gemm_bf16_tn(dh, op.wf_t, relu_mask=y3.view(B, -1), out_dtype=torch.bfloat16))
t('conv3 wgrad: dY^T', lambda: ops.transpose_bf16(dy3.view(-1, 64)))
t('conv3 wgrad: im2col_t', lambda: ops.im2col_t_bf16(x3, 3, 3, ones_row=True))
a3, b3 = ops.transpose_bf16(dy3.view(-1, 64)), ops.im2col_t_bf16(x3, 3, 3, ones_row=True)
t('conv3 wgrad: gemm', lambda: ops.gemm_bf16_tn(a3, b3))
t('conv3 dgrad: conv', lambda: ops.conv2d_nhwc_bf16(dy3, op.w3_flip, 3, 3, pad=(2, 2), relu_mask=x3))
t('conv2 wgrad: dY^T', lambda: ops.transpose_bf16(dy2.view(-1, 64)))
t('conv2 wgrad: im2col_t', lambda: ops.im2col_t_bf16(x2, 2, 2, ones_row=True))
a2, b2 = ops.transpose_bf16(dy2.view(-1, 64)), ops.im2col_t_bf16(x2, 2, 2, ones_row=True)
t('conv2 wgrad: gemm', lambda: ops.gemm_bf16_tn(a2, b2))
t('conv2 dgrad: conv', lambda: ops.conv2d_nhwc_bf16(dy2, op.w2_flip, 2, 2, pad=(1, 1), relu_mask=x2))
t('conv1 wgrad: dY^T', lambda: ops.transpose_bf16(dy1.view(-1, 32)))
t('conv1 wgrad: im2col_t', lambda: ops.im2col_t_bf16(x1, 2, 2, pixel_s2d=True, ones_row=True))
a1, b1 = ops.transpose_bf16(dy1.view(-1, 32)), ops.im2col_t_bf16(x1, 2, 2, pixel_s2d=True, ones_row=True)
t('conv1 wgrad: gemm', lambda: ops.gemm_bf16_tn(a1, b1))
t('conv3 wgrad: implicit (all)', lambda: ops.conv_wgrad_bf16(dy3.view(-1, 64), x3, 3, 3))
t('conv2 wgrad: implicit (all)', lambda: ops.conv_wgrad_bf16(dy2.view(-1, 64), x2, 2, 2))
t('conv1 wgrad: implicit (all)', lambda: ops.conv_wgrad_bf16(dy1.view(-1, 32), x1, 2, 2, s2d_order=True))
for name, us in rows:
    print(f'{name:28s} {us:9.0f} us')
print(f'{"total":28s} {sum(u for _, u in rows):9.0f} us')
