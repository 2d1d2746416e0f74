#!/usr/bin/env python
"""Whole-agent throughput at config C3's shape: `PPO.fit` on synthetic Atari frames, n_envs=256, n_steps=128, 4 epochs x 4
minibatches -- rollout (network inference, action sampling, environment step, writes into the time-major buffers) AND
update phase (GAE, gathers, network forward/backward, fused loss, clip+Adam), i.e. everything a training run does.

Variants: the device-resident batched environment (`SyntheticAtariDevice-v0`: no host round trip in the rollout) against a
Python list of 256 host environments (`SyntheticAtari-v0`: per-env loop + H2D of 7 MB of frames per step, what a gym
environment list costs), each with the Nature CNN on the tcgen05 kernels and in torch fp32.  Prints one markdown table and
one JSON line per variant.  Not the headline metric of bench.py (which excludes the network); it is the number a user of
`fit()` sees."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xagents_b200 import envs as xenvs, ops  # noqa: E402
from xagents_b200.agents import PPO, NatureCNN, NatureCnnTc, TorchModel  # noqa: E402

E, T, A = int(os.environ.get('N_ENVS', 256)), int(os.environ.get('N_STEPS', 128)), 6
STEPS = int(os.environ.get('TRAIN_STEPS', 3))


def run(env_id, tensor_cores, graph_rollout='auto'):
    torch.manual_seed(0)
    made = xenvs.create_envs(env_id, E, preprocess=True, device='cuda:0')
    net = TorchModel((NatureCnnTc if tensor_cores else NatureCNN)(4, A).cuda())
    agent = PPO(made, net, n_steps=T, mini_batches=4, ppo_epochs=4, quiet=True, seed=1)
    agent.graph_rollout = graph_rollout
    phases = {'rollout': 0.0, 'update': 0.0}
    inner_batch, inner_epochs = agent.get_batch, agent.run_ppo_epochs

    def timed(name, fn):
        def wrapper(*args):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = fn(*args)
            torch.cuda.synchronize()
            phases[name] += time.perf_counter() - t0
            return out
        return wrapper

    agent.get_batch, agent.run_ppo_epochs = timed('rollout', inner_batch), timed('update', inner_epochs)
    agent.fit(max_steps=3 * T * E)                                 # warm-up: first launches, graph capture, allocator (the permutation draw
    #                                                                of the THIRD update still takes one cudaMalloc, 0.5-80 ms: scripts/fit_step_times.py)
    phases.update(rollout=0.0, update=0.0)
    ops.reset_launch_count()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    agent.fit(max_steps=agent.steps + STEPS * T * E)
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / STEPS
    return {'env': env_id, 'network': 'tcgen05 kernels (bf16 operands, fp32 accumulate)' if tensor_cores else 'torch fp32',
            'rollout': 'one CUDA graph' if agent._rollout_graph else 'eager loop', 'n_envs': E, 'n_steps': T, 'train_steps_timed': STEPS, 'ms_per_train_step': sec * 1e3,
            'rollout_ms': phases['rollout'] / STEPS * 1e3, 'update_ms': phases['update'] / STEPS * 1e3,
            'env_steps_per_sec': T * E / sec, 'hot_path_kernel_launches_per_train_step': ops.launch_count() // STEPS,
            'finite': bool(torch.isfinite(net.flat_param).all())}


rows = [run('SyntheticAtariDevice-v0', True), run('SyntheticAtariDevice-v0', True, graph_rollout=False), run('SyntheticAtariDevice-v0', False),
        run('SyntheticAtari-v0', True)]
print('| environments | network | rollout | ms / train step | rollout ms | update ms | env-steps/s (whole training) |')
print('|---|---|---|---|---|---|---|')
for r in rows:
    print(f"| {r['env']} x {r['n_envs']} | {r['network']} | {r['rollout']} | {r['ms_per_train_step']:.1f} | {r['rollout_ms']:.1f} | {r['update_ms']:.1f} | "
          f"{r['env_steps_per_sec'] / 1e3:.0f} k |")
for r in rows:
    print(json.dumps(r))
