#!/usr/bin/env python
"""BASELINE configs C1 (PPO, CartPole-shaped, n_envs=16, n_steps=128) and C2 (A2C, Pong-shaped frames, n_envs=16,
n_steps=5): latency-bound shapes.  Reports microseconds per train step of the hot path (eager launches and one
CUDA-graph replay) next to the oracle's CPU time for the same step.  Not roofline-meaningful (a few hundred KB)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from xagents_b200 import synthetic  # noqa: E402
from xagents_b200.hotpath import A2CHotPath, PPOHotPath  # noqa: E402

dev = 'cuda:0'


def timeit(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


print('| config | launches/step | eager us/step | graph replay us/step | CPU oracle us/step | speed-up (graph) |')
print('|---|---|---|---|---|---|')
# C1
T, E, A = 128, 16, 2
ro = synthetic.make_rollout(T, E, obs_shape=(4,), obs_dtype='float32', n_actions=A, epochs=4, p_done=0.02)
hp = PPOHotPath(T, E, (4,), A, obs_dtype=torch.float32, device=dev)
hp.load(ro)
hp.perms.copy_(torch.as_tensor(np.stack(ro.permutations)))
hp.actor_out.normal_()
hp.critic_out.normal_()
hp.prepare()
t_eager = timeit(hp.run)
replay = hp.capture()
t_graph = timeit(replay)
t0 = time.perf_counter()
for _ in range(5):
    oracle.ppo_train_step(ro.obs, ro.rewards, ro.dones, ro.values, ro.last_values, ro.actions, ro.log_probs, ro.permutations,
                          ro.new_logits, ro.new_values, mini_batches=4)
t_cpu = (time.perf_counter() - t0) / 5 * 1e6
print(f'| C1 PPO CartPole-shaped T=128 E=16 (4x4 minibatches) | {hp.kernel_launches_per_step} | {t_eager:.0f} | {t_graph:.0f} | {t_cpu:.0f} | {t_cpu / t_graph:.0f}x |')
# C2
T, E, A = 5, 16, 6
ro = synthetic.make_rollout(T, E, n_actions=A, epochs=0)
a2c = A2CHotPath(T, E, A, device=dev)
for k in ('rewards', 'values', 'last_values', 'dones', 'actions'):
    getattr(a2c, k).copy_(torch.as_tensor(getattr(ro, k)))
a2c.actor_out.normal_()
a2c.critic_out.normal_()
a2c.prepare()
t_eager = timeit(a2c.run)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(dev)
with torch.cuda.stream(s):
    a2c.prepare(s)
    a2c.run()
    s.synchronize()
    with torch.cuda.graph(g, stream=s):
        a2c.run()
t_graph = timeit(g.replay)
logits = ro.new_logits
t0 = time.perf_counter()
for _ in range(50):
    ret = oracle.nstep_returns(ro.rewards, ro.dones, ro.last_values, 0.99)
    flat = oracle.concat_step_batches(ro.obs, ret, ro.actions, ro.values)
    logp, ent, _ = oracle.categorical_logp_entropy(logits, flat[2])
    oracle.a2c_loss(logp, ro.new_values, ent, flat[3], flat[1], 0.01, 0.5)
    oracle.a2c_loss_grads(logits, ro.new_values, flat[2], flat[3], flat[1], 0.01, 0.5)
t_cpu = (time.perf_counter() - t0) / 50 * 1e6
print(f'| C2 A2C Pong-shaped T=5 E=16 (84x84x4 uint8 frames, no reorder copy) | 2 | {t_eager:.0f} | {t_graph:.0f} | {t_cpu:.0f} | {t_cpu / t_graph:.0f}x |')
