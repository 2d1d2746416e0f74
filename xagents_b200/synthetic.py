"""Seeded synthetic rollouts in the shapes of BASELINE.json's configs.

The same NumPy arrays feed the CUDA path and the CPU oracle.  Rollout fields are time-major, exactly
what A2C.get_batch / PPO.get_batch hold after `np.asarray(..., np.float32)` (xagents/ppo/agent.py:202-210),
except that image observations stay uint8 (the reference stores the same integers as fp32).
"""
from dataclasses import dataclass, field

import numpy as np

DATA_SEED = 1234
PERM_SEED = 4321

# name -> (n_steps, n_envs, obs_shape, obs_dtype, n_actions)
WORKLOADS = {
    'c1_ppo_cartpole': (128, 16, (4,), 'float32', 2),
    'c2_a2c_pong': (5, 16, (84, 84, 4), 'uint8', 6),
    'c3_ppo_atari_e256': (128, 256, (84, 84, 4), 'uint8', 6),
    'c4_ppo_atari_e4096': (128, 4096, (84, 84, 4), 'uint8', 6),
}


@dataclass
class Rollout:
    n_steps: int
    n_envs: int
    n_actions: int
    obs: np.ndarray              # [T, E, *obs_shape]
    rewards: np.ndarray          # [T, E] fp32
    values: np.ndarray           # [T, E] fp32   (critic output during the rollout)
    last_values: np.ndarray      # [E] fp32      (bootstrap V(s_T), ppo/agent.py:72-79)
    dones: np.ndarray            # [T+1, E] fp32 (row t+1 gates step t)
    actions: np.ndarray          # [T, E] fp32-encoded action ids
    log_probs: np.ndarray        # [T, E] fp32   (old log-probs)
    new_logits: np.ndarray       # [N, A] fp32, env-major: stand-in for the forward pass during updates
    new_values: np.ndarray       # [N] fp32, env-major
    permutations: list = field(default_factory=list)   # K int32 [N] arrays of env-major sample ids

    @property
    def batch_size(self):
        return self.n_steps * self.n_envs


def _log_softmax(x):
    z = x - x.max(axis=-1, keepdims=True)
    return z - np.log(np.exp(z).sum(axis=-1, keepdims=True))


def make_rollout(n_steps, n_envs, obs_shape=(84, 84, 4), obs_dtype='uint8', n_actions=6, epochs=4, p_done=0.01,
                 seed=DATA_SEED, perm_seed=PERM_SEED, with_obs=True, sparse_rewards=False):
    rng = np.random.default_rng(seed)
    T, E, A = n_steps, n_envs, n_actions
    N = T * E
    if not with_obs:
        obs = np.zeros((T, E, 0), np.uint8)
    elif obs_dtype == 'uint8':
        obs = rng.integers(0, 256, size=(T, E) + tuple(obs_shape), dtype=np.uint8)
    else:
        obs = rng.standard_normal((T, E) + tuple(obs_shape)).astype(np.float32)
    if sparse_rewards:
        rewards = (np.sign(rng.standard_normal((T, E))) * (rng.random((T, E)) < 0.02)).astype(np.float32)
    else:
        rewards = rng.standard_normal((T, E)).astype(np.float32)
    all_values = rng.standard_normal((T + 1, E)).astype(np.float32)
    dones = (rng.random((T + 1, E)) < p_done).astype(np.float32)
    actions = rng.integers(0, A, size=(T, E)).astype(np.float32)
    old_logits = rng.standard_normal((T, E, A)).astype(np.float32)
    lsm = _log_softmax(old_logits)
    log_probs = np.take_along_axis(lsm, actions.astype(np.int64)[..., None], axis=-1)[..., 0].astype(np.float32)
    # env-major copies for the "model outputs during the update" stand-ins
    flat_logits = old_logits.swapaxes(0, 1).reshape(N, A)
    flat_values = all_values[:-1].swapaxes(0, 1).reshape(N)
    new_logits = (flat_logits + 0.1 * rng.standard_normal((N, A))).astype(np.float32)
    new_values = (flat_values + 0.1 * rng.standard_normal(N)).astype(np.float32)
    prng = np.random.default_rng(perm_seed)
    perms = [prng.permutation(N).astype(np.int32) for _ in range(epochs)]
    return Rollout(T, E, A, obs, rewards, all_values[:-1].copy(), all_values[-1].copy(), dones, actions, log_probs,
                   new_logits, new_values, perms)


def make_workload(name, **overrides):
    T, E, shape, dtype, A = WORKLOADS[name]
    kw = dict(n_steps=T, n_envs=E, obs_shape=shape, obs_dtype=dtype, n_actions=A)
    kw.update(overrides)
    return make_rollout(**kw)
