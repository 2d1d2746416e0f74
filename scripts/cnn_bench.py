#!/usr/bin/env python
"""Nature CNN forward: tcgen05 pipeline (NatureCnnTcForward) vs torch/cuDNN in fp32(TF32 off), bf16 autocast and channels_last."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xagents_b200 import ops  # noqa: E402
from xagents_b200.agents import NatureCNN  # noqa: E402
from xagents_b200.agents.tc_conv import NatureCnnTcForward  # noqa: E402

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


net = NatureCNN(4, 6).cuda()
tc = NatureCnnTcForward(net)
FLOP = 18.7e6 * 2 / 2          # 18.7 MFLOP forward per sample
print('| batch | ours us | ours TFLOP/s | torch fp32 us | torch bf16 autocast us | bf16 channels_last us |')
print('|---|---|---|---|---|---|')
for B in (256, 2048, 8192):
    x = torch.randint(0, 256, (B, 84, 84, 4), dtype=torch.uint8, device=dev)
    t_ours = timeit(lambda: tc(x))
    with torch.no_grad():
        xf = x.float() / 255.0
        t_fp32 = timeit(lambda: net(xf))
        with torch.autocast('cuda', dtype=torch.bfloat16):
            t_bf16 = timeit(lambda: net(xf))
        net_cl = net.to(memory_format=torch.channels_last)
        with torch.autocast('cuda', dtype=torch.bfloat16):
            t_cl = timeit(lambda: net_cl(xf))
    print(f'| {B} | {t_ours:.0f} | {18.7e6 * B / t_ours / 1e6:.0f} | {t_fp32:.0f} | {t_bf16:.0f} | {t_cl:.0f} |')
# per layer (B = 8192)
B = 8192
x = torch.randint(0, 256, (B, 84, 84, 4), dtype=torch.uint8, device=dev)
x1 = ops.space_to_depth_u8_bf16(x, 4)
t0 = timeit(lambda: ops.space_to_depth_u8_bf16(x, 4))
x2 = ops.conv2d_nhwc_bf16(x1, tc.w1, 2, 2, bias=tc.b1, relu=True, out_s2d=True)
t1 = timeit(lambda: ops.conv2d_nhwc_bf16(x1, tc.w1, 2, 2, bias=tc.b1, relu=True, out_s2d=True))
x3 = ops.conv2d_nhwc_bf16(x2, tc.w2, 2, 2, bias=tc.b2, relu=True)
t2 = timeit(lambda: ops.conv2d_nhwc_bf16(x2, tc.w2, 2, 2, bias=tc.b2, relu=True))
x4 = ops.conv2d_nhwc_bf16(x3, tc.w3, 3, 3, bias=tc.b3, relu=True)
t3 = timeit(lambda: ops.conv2d_nhwc_bf16(x3, tc.w3, 3, 3, bias=tc.b3, relu=True))
t4 = timeit(lambda: ops.gemm_bf16_tn(x4.view(B, -1), tc.wf, bias=tc.bf_, relu=True, out_dtype=torch.bfloat16))
print(f'\nper layer at B={B} (us): space-to-depth {t0:.0f}, conv1 {t1:.0f} ({2*B*400*256*32/t1/1e6:.0f} TFLOP/s), '
      f'conv2 {t2:.0f} ({2*B*81*512*64/t2/1e6:.0f}), conv3 {t3:.0f} ({2*B*49*576*64/t3/1e6:.0f}), fc {t4:.0f} ({2*B*3136*512/t4/1e6:.0f})')

# ---- training step: forward + backward of the whole network (B = 8192, the C3 minibatch)
from xagents_b200.agents import NatureCnnTc  # noqa: E402

B = 8192
x = torch.randint(0, 256, (B, 84, 84, 4), dtype=torch.uint8, device=dev)
da, dc = torch.randn((B, 6), device=dev) / B, torch.randn(B, device=dev) / B
tcn = NatureCnnTc(4, 6).cuda()


def ours_step():
    a, c = tcn(x)
    torch.autograd.backward([a, c], [da, dc])


def torch_step(autocast):
    with torch.autocast('cuda', dtype=torch.bfloat16, enabled=autocast):
        a, c = net(x.float() / 255.0)
    torch.autograd.backward([a.float(), c.float().reshape(-1)], [da, dc])


t_ours = timeit(ours_step, 10)
t_fp32 = timeit(lambda: torch_step(False), 5)
t_bf16 = timeit(lambda: torch_step(True), 5)
fl = 3 * 18.7e6 * B
print(f'\nforward+backward at B={B}: ours {t_ours:.0f} us ({fl / t_ours / 1e6:.0f} TFLOP/s), torch fp32 {t_fp32:.0f} us, '
      f'torch bf16 autocast {t_bf16:.0f} us  -> {t_bf16 / t_ours:.1f}x / {t_fp32 / t_ours:.1f}x')

# ---- the same step through the native plan (two C calls per minibatch, agents/tc_plan.py): what TorchModel runs
from xagents_b200.agents import TorchModel  # noqa: E402

tm = TorchModel(NatureCnnTc(4, 6).cuda())
plan = tm.module.plan(B, tm.flat_param, tm.flat_grad.numel())


def plan_step():
    plan.forward(x)
    plan.backward(da, dc, tm.flat_grad)


t_plan = timeit(plan_step, 20)
t_fwd = timeit(lambda: plan.forward(x), 20)
x_s2d = ops.space_to_depth_u8_bf16(x, 4)


def plan_step_s2d():
    plan.forward(x_s2d)
    plan.backward(da, dc, tm.flat_grad)


t_plan_s2d = timeit(plan_step_s2d, 20)
print(f'native plan at B={B}: forward+backward {t_plan:.0f} us ({fl / t_plan / 1e6:.0f} TFLOP/s), forward alone {t_fwd:.0f} us, '
      f'forward+backward from space-to-depth input (the fused gather\'s output) {t_plan_s2d:.0f} us')
t_u8 = timeit(lambda: ops.conv2d_u8_s2d_bf16(x, tc.w1, 2, 2, bias=tc.b1, relu=True, out_s2d=True), 20)
print(f'first layer straight from uint8 frames at B={B}: {t_u8:.0f} us (space-to-depth {t0:.0f} + conv1 {t1:.0f} = {t0 + t1:.0f} us before); '
      f'{B * (28224 + 400 * 32 * 2) / t_u8 / 1e3:.0f} GB/s of frames + output')
x1_out = torch.empty((B, 21, 21, 64), dtype=torch.bfloat16, device=dev)
t_u8s = timeit(lambda: ops.conv2d_u8_s2d_bf16(x, tc.w1, 2, 2, bias=tc.b1, relu=True, out_s2d=True, x_s2d_out=x1_out), 20)
print(f'... and also storing the bf16 space-to-depth tensor for the backward pass: {t_u8s:.0f} us')
# ... reading the minibatch through a permutation of a 32768-frame rollout (the gather folded into the layer)
store = torch.randint(0, 256, (128 * 256, 84, 84, 4), dtype=torch.uint8, device=dev)
perm = torch.randperm(128 * 256, device=dev)[:B].to(torch.int32)
t_idx = timeit(lambda: ops.conv2d_u8_s2d_bf16(store, tc.w1, 2, 2, bias=tc.b1, relu=True, out_s2d=True, x_s2d_out=x1_out, idx=perm, time_major=(128, 256)), 20)
mb = torch.empty((B, 84, 84, 4), dtype=torch.uint8, device=dev)
t_gather = timeit(lambda: ops.gather_rows(store.view(128, 256, 84, 84, 4), perm, out=mb, time_major=(128, 256)), 20)
print(f'... through a permutation of a 32768-frame rollout (frames fetched by id, no gathered copy): {t_idx:.0f} us; '
      f'the gather it replaces: {t_gather:.0f} us + {t_u8s:.0f} us')
