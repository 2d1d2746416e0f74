"""The reference's CPU path for one PPO train step, timed (bench.py cpu_baseline / --impl reference).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  `kind: "port"`: TensorFlow is not installable in
this image, so this is the restatement, not the literal reference: NumPy for what the reference does in
NumPy (np.asarray of the rollout to fp32, the GAE loop, concat_step_batches -- xagents/ppo/agent.py:202-213,
80-94, base.py:559-564) and torch-CPU ops, on all host threads, standing in 1:1 for the TF-CPU ops
(tf.gather x5 per minibatch with ALL K*M minibatches materialised up front as get_mini_batches does,
reduce_mean/reduce_std, the loss and its tape gradient -- ppo/agent.py:139-155, 180-183, 112-134).
Observations are fp32 as in the reference (`layout='reference'`) or uint8 (`layout='uint8'`).
The model forward/backward is not part of the metric on either arm: per-minibatch model outputs are inputs.
"""
import time

import numpy as np
import torch

from . import hotpath, torch_ref


def ppo_train_step_cpu(ro, mini_batches=4, layout='reference', gamma=0.99, lam=0.95, clip_norm=0.1, entropy_coef=0.01,
                       value_loss_coef=0.5, advantage_epsilon=1e-8):
    """Run one train step of the path on the host; returns (seconds, last loss)."""
    T, E = ro.n_steps, ro.n_envs
    N = T * E
    B = N // mini_batches
    t0 = time.perf_counter()
    # PPO.get_batch: lists -> fp32 arrays (ppo/agent.py:202-210)
    states = np.asarray(ro.obs, np.float32) if layout == 'reference' else np.asarray(ro.obs)
    returns = hotpath.gae_returns(ro.rewards, ro.dones, ro.values, ro.last_values, gamma, lam)
    flat_states, flat_actions, flat_returns, flat_values, flat_logp = hotpath.concat_step_batches(
        states, ro.actions, returns, ro.values, ro.log_probs)
    fields = [torch.from_numpy(x) for x in (flat_states, flat_actions, flat_returns, flat_values, flat_logp)]
    new_logits, new_values = torch.from_numpy(ro.new_logits), torch.from_numpy(ro.new_values)
    # get_mini_batches: every minibatch of every epoch gathered before the first update (ppo/agent.py:149-155)
    minibatches = []
    for perm in ro.permutations:
        perm_t = torch.from_numpy(np.asarray(perm)).long()
        for lo, hi in hotpath.minibatch_slices(N, B):
            idx = perm_t[lo:hi]
            minibatches.append((idx, [torch_ref.gather_rows(f, idx) for f in fields]))
    loss = None
    for idx, (s_mb, a_mb, r_mb, v_mb, lp_mb) in minibatches:
        adv = torch_ref.normalize_advantages(r_mb, v_mb, advantage_epsilon)
        loss, _, _ = torch_ref.ppo_loss_fwd_bwd(new_logits.index_select(0, idx), new_values.index_select(0, idx), a_mb,
                                                v_mb, r_mb, lp_mb, adv, clip_norm, entropy_coef, value_loss_coef)
    return time.perf_counter() - t0, float(loss)


def a2c_train_step_cpu(ro, layout='reference', gamma=0.99, entropy_coef=0.01, value_loss_coef=0.5):
    """A2C.train_step on the host (xagents/a2c/agent.py:173-218): np.asarray of the rollout, the n-step returns loop,
    concat_step_batches (the env-major flatten COPY of states, returns, actions, values), then the loss and its tape
    gradient over all T*E samples.  Model outputs are inputs, as on the GPU arm."""
    t0 = time.perf_counter()
    states = np.asarray(ro.obs, np.float32) if layout == 'reference' else np.asarray(ro.obs)
    returns = hotpath.nstep_returns(ro.rewards, ro.dones, ro.last_values, gamma)
    flat_states, flat_returns, flat_actions, flat_values = hotpath.concat_step_batches(states, returns, ro.actions, ro.values)
    loss, _, _ = torch_ref.a2c_loss_autograd(ro.new_logits, ro.new_values, flat_actions, flat_values, flat_returns, entropy_coef,
                                             value_loss_coef)
    return time.perf_counter() - t0, float(loss['loss'])


def time_cpu_baseline(ro, steps=1, warmup=0, algo='ppo', **kw):
    """Median seconds per train step over `steps` runs (after `warmup`), env-steps/s, threads used."""
    step = a2c_train_step_cpu if algo == 'a2c' else ppo_train_step_cpu
    for _ in range(warmup):
        step(ro, **kw)
    times = [step(ro, **kw)[0] for _ in range(steps)]
    sec = float(np.median(times))
    return dict(seconds_per_step=sec, env_steps_per_sec=ro.n_steps * ro.n_envs / sec, threads=torch.get_num_threads(),
                times=times)
