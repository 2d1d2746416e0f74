#!/usr/bin/env python
"""Per-train-step wall times of PPO.fit at C3's shape (rollout as one CUDA graph + update phase), to look for outliers:
python scripts/fit_step_times.py [steps]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xagents_b200 import envs as xenvs  # noqa: E402
from xagents_b200.agents import PPO, NatureCnnTc, TorchModel  # noqa: E402

E, T, A = 256, 128, 6
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
torch.manual_seed(0)
made = xenvs.create_envs('SyntheticAtariDevice-v0', E, preprocess=True, device='cuda:0')
net = TorchModel(NatureCnnTc(4, A).cuda())
agent = PPO(made, net, n_steps=T, mini_batches=4, ppo_epochs=4, quiet=True, seed=1)
agent.fit(max_steps=T * E)
torch.cuda.synchronize()
rows = []
acc = {}


def timed(obj, name):
    fn = getattr(obj, name)

    def wrapper(*a, **kw):
        t = time.perf_counter()
        out = fn(*a, **kw)
        acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t) * 1e3
        return out
    setattr(obj, name, wrapper)


timed(agent, '_next_permutation')
timed(agent, 'hot_path')
timed(net, 'forward_into')
timed(net, 'backward_and_step')
for k in range(steps):
    acc.clear()
    st0 = torch.cuda.memory_stats()
    t0 = time.perf_counter()
    batch = agent.get_batch()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    agent.run_ppo_epochs(*batch)
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    st1 = torch.cuda.memory_stats()
    rows.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t1) * 1e3, dict(acc),
                 st1['num_device_alloc'] - st0['num_device_alloc'], st1['num_device_free'] - st0['num_device_free'], st1['num_alloc_retries'] - st0['num_alloc_retries']))
print('step: rollout ms | update: host issue ms, until done ms | host ms inside: permutation, hot_path, forward_into, backward_and_step | cudaMalloc, cudaFree, retries')
for k, (r, h, u, a, na, nf, nr) in enumerate(rows):
    print(f'{k:3d}: {r:7.2f} | {h:7.2f} {u:7.2f} | ' + ' '.join(f"{a.get(n, 0.0):6.2f}" for n in ('_next_permutation', 'hot_path', 'forward_into', 'backward_and_step')) + f' | {na} {nf} {nr}')
