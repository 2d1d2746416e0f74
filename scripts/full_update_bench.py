#!/usr/bin/env python
"""The whole PPO update phase at C3 (n_envs=256, n_steps=128, 4 epochs x 4 minibatches of 8192 frames): GAE, then per
minibatch gather -> network forward -> fused loss -> network backward -> fused clip+Adam, i.e. rows (a)-(f) of the scope
table working together.  Model: the Nature CNN on the tcgen05 kernels (NatureCnnTc) vs the same network in torch fp32
(the reference's dtype; cuDNN/cuBLAS) and torch bf16 autocast.  Everything else (hot-path kernels) is identical."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xagents_b200 import ops  # noqa: E402
from xagents_b200.agents import NatureCNN, NatureCnnTc, TorchModel  # noqa: E402

dev = 'cuda:0'
T, E, A, K, M = 128, 256, 6, 4, 4
N, B = T * E, T * E // 4
g = torch.Generator(device=dev)
g.manual_seed(0)
obs = torch.randint(0, 256, (T, E, 84, 84, 4), dtype=torch.uint8, device=dev, generator=g)
rewards, values = torch.randn((T, E), device=dev, generator=g), torch.randn((T, E), device=dev, generator=g)
last_values = torch.randn(E, device=dev, generator=g)
dones = (torch.rand((T + 1, E), device=dev, generator=g) < 0.01).float()
actions = torch.randint(0, A, (T, E), device=dev, generator=g).float()
log_probs = -torch.rand((T, E), device=dev, generator=g) - 0.5
perms = torch.stack([torch.randperm(N, device=dev, generator=g).to(torch.int32) for _ in range(K)])
workspace = ops.loss_workspace(B, dev)


class Autocast:
    def __init__(self, inner):
        self.inner = inner
        self.__dict__.update({k: getattr(inner, k) for k in ('output_is_softmax', 'comm')})

    def forward(self, x, training=True):
        with torch.autocast('cuda', dtype=torch.bfloat16):
            a, c = self.inner.forward(x, training)
        return a.float(), c.float()

    def backward_and_step(self, *args):
        a, c = self.inner._outputs
        self.inner._outputs = (a.float(), c.float())
        return self.inner.backward_and_step(*args)


def update_phase(net, fused_gather=False):
    ret = ops.gae_returns(rewards, values, last_values, dones, 0.99, 0.95)
    offsets = [k * N + m * B for k in range(K) for m in range(M)] + [K * N]
    moments = ops.adv_moments(ret, values, perms.view(-1), offsets, time_major=(T, E))
    for k in range(K):
        for m in range(M):
            idx = perms[k, m * B:(m + 1) * B]
            states = (ops.gather_s2d_u8_bf16(obs, idx, time_major=(T, E)) if fused_gather
                      else ops.gather_rows(obs, idx, time_major=(T, E)))
            actor, critic = net.forward(states, training=True)
            _, d_actor, d_values, _ = ops.ppo_loss(actor, critic, actions, log_probs, values, ret, idx=idx, time_major=(T, E),
                                                   moments=moments[k * M + m], workspace=workspace)
            net.backward_and_step(d_actor, d_values, 0.5)


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


torch.manual_seed(0)
rows = []
for name, net in (('Nature CNN on tcgen05 kernels, gather fused with /255 + space-to-depth', TorchModel(NatureCnnTc(4, A).cuda())),
                  ('Nature CNN on tcgen05 kernels (bf16 operands, fp32 accumulate)', TorchModel(NatureCnnTc(4, A).cuda())),
                  ('torch fp32 (cuDNN / cuBLAS)', TorchModel(NatureCNN(4, A).cuda())),
                  ('torch bf16 autocast', Autocast(TorchModel(NatureCNN(4, A).cuda())))):
    ms = timeit(lambda: update_phase(net, fused_gather='fused' in name))
    rows.append((name, ms))
print('| network path | ms per update phase (16 minibatches of 8192) | env-steps/s through the update phase |')
print('|---|---|---|')
for name, ms in rows:
    print(f'| {name} | {ms:.1f} | {N / ms * 1e3 / 1e3:.0f} k |')
