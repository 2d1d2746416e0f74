"""Environments for the CLI / factory (`create_envs`, xagents/utils/common.py:145-167).

The reference builds its environments with `gym.make`; neither gym nor ALE exists in this image, so the two
environments BASELINE.json's configs name ship here in the gym API the reference's agents use
(`reset() -> state`, `step(action) -> (state, reward, done, info)`, `observation_space.shape`,
`action_space.n`, `spec.id`):

* `CartPole-v1` -- the classic cart-pole balance task (Barto, Sutton & Anderson 1983) with the standard
  constants (Euler integration at 0.02 s, 12 degrees / 2.4 m limits, 500-step episodes, reward 1 per step):
  config C1 runs end to end and learns.
* `SyntheticAtariDevice-v0`, `CartPoleDevice-v1` -- the same two tasks as ONE device-resident batched environment each
  (`BatchedSyntheticAtari`, `BatchedCartPole`): rollouts without host round trips.
* `SyntheticAtari-v0` (aliases `SyntheticPong-v0`, and any `*NoFrameskip-v4` id when ALE is absent and
  `XAGENTS_B200_SYNTHETIC_ATARI=1`) -- 84x84xC uint8 frames drawn i.i.d., sparse +-1 rewards, geometric
  episode lengths: the "synthetic Atari frames" of configs C2-C4 behind the env interface.

Any other id is handed to `gym` / `gymnasium` if one is importable (gymnasium's 5-tuple step is folded back to
the 4-tuple the agents read); otherwise `create_envs` fails loudly.
"""
import math
import os

import numpy as np


class Discrete:
    def __init__(self, n, seed=None):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.int64
        self._rng = np.random.default_rng(seed)

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        return int(self._rng.integers(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n

    def __repr__(self):
        return f'Discrete({self.n})'


class Box:
    def __init__(self, low, high, shape, dtype=np.float32, seed=None):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)
        self._shape = self.shape
        self._rng = np.random.default_rng(seed)

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        if self.dtype == np.uint8:
            return self._rng.integers(0, 256, self.shape, dtype=np.uint8)
        return self._rng.uniform(np.maximum(self.low, -1e3), np.minimum(self.high, 1e3), self.shape).astype(self.dtype)

    def __repr__(self):
        return f'Box{self.shape}'


class _Spec:
    def __init__(self, env_id):
        self.id = env_id


class CartPole:
    """Cart-pole balance, gym `CartPole-v1` dynamics and termination rules."""
    GRAVITY, CART_MASS, POLE_MASS, HALF_LENGTH, FORCE, TAU = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02
    THETA_LIMIT, X_LIMIT, MAX_STEPS = 12 * 2 * math.pi / 360, 2.4, 500

    def __init__(self, env_id='CartPole-v1', seed=None):
        self.spec = _Spec(env_id)
        high = np.array([2 * self.X_LIMIT, np.finfo(np.float32).max, 2 * self.THETA_LIMIT, np.finfo(np.float32).max], np.float32)
        self.observation_space = Box(-high, high, (4,), np.float32)
        self.action_space = Discrete(2)
        self._rng = np.random.default_rng(seed)
        self.state, self.t = None, 0

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def reset(self, **kwargs):
        self.state = self._rng.uniform(-0.05, 0.05, 4)
        self.t = 0
        return self.state.astype(np.float32)

    def step(self, action):
        assert self.state is not None, 'Cannot call env.step() before calling reset()'
        x, x_dot, th, th_dot = self.state
        force = self.FORCE if int(action) == 1 else -self.FORCE
        cos, sin = math.cos(th), math.sin(th)
        total = self.CART_MASS + self.POLE_MASS
        pml = self.POLE_MASS * self.HALF_LENGTH
        temp = (force + pml * th_dot * th_dot * sin) / total
        th_acc = (self.GRAVITY * sin - cos * temp) / (self.HALF_LENGTH * (4.0 / 3.0 - self.POLE_MASS * cos * cos / total))
        x_acc = temp - pml * th_acc * cos / total
        self.state = np.array([x + self.TAU * x_dot, x_dot + self.TAU * x_acc, th + self.TAU * th_dot, th_dot + self.TAU * th_acc])
        self.t += 1
        fell = abs(self.state[0]) > self.X_LIMIT or abs(self.state[2]) > self.THETA_LIMIT
        done = bool(fell or self.t >= self.MAX_STEPS)
        return self.state.astype(np.float32), 1.0, done, {'TimeLimit.truncated': bool(not fell and done)}

    def close(self):
        pass


class SyntheticAtari:
    """Atari-shaped frames without an emulator: i.i.d. uint8 pixels, +-1 rewards w.p. `p_reward`, episode ends w.p.
    `p_done` per step (SURVEY.md 8d's synthetic distribution, served through the env interface)."""

    def __init__(self, env_id='SyntheticAtari-v0', channels=4, n_actions=6, p_done=0.01, p_reward=0.02, seed=None, pool=64):
        self.spec = _Spec(env_id)
        self.observation_space = Box(0, 255, (84, 84, channels), np.uint8)
        self.action_space = Discrete(n_actions)
        self.p_done, self.p_reward = p_done, p_reward
        self._pool_size = pool
        self.seed(seed)

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        # frames are drawn from a pool generated once: drawing 28 KB of fresh random bytes per step and env would make
        # the generator, not the agent, the thing a training run measures
        self._pool = self._rng.integers(0, 256, (self._pool_size,) + self.observation_space.shape, dtype=np.uint8)
        return [seed]

    def _frame(self):
        return self._pool[int(self._rng.integers(self._pool_size))]

    def reset(self, **kwargs):
        return self._frame()

    def step(self, action):
        u = self._rng.random(3)
        reward = 0.0 if u[0] >= self.p_reward else (1.0 if u[1] < 0.5 else -1.0)
        return self._frame(), reward, bool(u[2] < self.p_done), {}

    def close(self):
        pass


class BatchedSyntheticAtari:
    """All `n` synthetic-frame environments as ONE object whose state lives on the device (SURVEY.md 8f-3 taken to the
    environment side): `reset_all()` / `step_all(actions)` take and return device tensors, so a rollout step is
    network -> sampler -> this, with no host round trip, no per-environment Python loop and no H2D copy of frames
    (256 x 28 KB per step at config C3).  Same distribution as `SyntheticAtari` (i.i.d. uint8 frames from a pool,
    +-1 rewards w.p. `p_reward`, episode ends w.p. `p_done`), same reset semantics as BaseAgent.step_envs
    (xagents/base.py:408-426): `step_all` returns the terminal frame of a finished episode while `states` already holds
    the frame after the reset.  Looks like a sequence of `n` environments to code that reads `len(envs)` or
    `envs[0].observation_space`; agents detect `batched` and call the two batch methods instead of looping."""
    batched = True

    def __init__(self, n, env_id='SyntheticAtariDevice-v0', channels=4, n_actions=6, p_done=0.01, p_reward=0.02, seed=None,
                 pool=256, device='cuda:0'):
        import torch
        self.torch, self.n, self.device = torch, int(n), torch.device(device)
        self.spec = _Spec(env_id)
        self.observation_space = Box(0, 255, (84, 84, channels), np.uint8)
        self.action_space = Discrete(n_actions)
        self.p_done, self.p_reward, self._pool_size = p_done, p_reward, pool
        self.states = None
        self.seed(seed)

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return self

    def __iter__(self):
        raise TypeError('a batched environment is not a list of environments: call reset_all() / step_all(actions)')

    def seed(self, seed=None):
        torch = self.torch
        self._seed = int(seed) if seed is not None else int(np.random.SeedSequence().entropy % (2 ** 63))
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(self._seed)
        self._pool = torch.randint(0, 256, (self._pool_size,) + self.observation_space.shape, dtype=torch.uint8,
                                   device=self.device, generator=self._gen)
        # On the GPU a step is ONE kernel (csrc/synth_env.cu) drawing from Philox at (seed, *counter + host offset): the
        # counter lives in device memory so that a captured rollout draws fresh numbers on every replay.
        self.native_step = self.device.type == 'cuda'
        if self.native_step:
            self._counter = torch.zeros(1, dtype=torch.int64, device=self.device)
            self._host_offset = 0
        return [seed]

    def _frames(self):
        idx = self.torch.randint(0, self._pool_size, (self.n,), device=self.device, generator=self._gen)
        return self._pool.index_select(0, idx)

    def reset_all(self):
        self.states = self._frames()
        return self.states

    # everything `step_all` reads and replaces, by attribute name: what a captured rollout (CUDA graph) must find at fixed
    # addresses on entry and what it hands back on exit
    GRAPH_STATE = ('states',)

    # ---- random streams around a captured rollout (agents/a2c.py:_capture_rollout) ----
    def rng_state(self):
        state = {'gen': self._gen.get_state()}
        if self.native_step:
            state['counter'], state['host_offset'] = int(self._counter.item()), self._host_offset
        return state

    def set_rng_state(self, state):
        self._gen.set_state(state['gen'])
        if self.native_step:
            self._counter.fill_(state['counter'])
            self._host_offset = state['host_offset']

    def sync_rng(self):
        """Fold the host-side step count into the device counter, in stream order: before a capture (the captured steps then
        count from the counter alone) and as the captured rollout's last node (every replay advances the stream)."""
        if self.native_step and self._host_offset:
            from . import ops
            ops.bump_u64(self._counter, self._host_offset)
            self._host_offset = 0

    def register_graph(self, graph):
        if not self.native_step:
            graph.register_generator_state(self._gen)

    def step_into(self, new_states, rewards, dones, episode_sums=None, sums_log=None):
        """The native step writing into the caller's tensors (rows of the rollout buffers); `self.states` is updated in place."""
        from . import ops
        ops.synth_env_step(self._pool, self.states, new_states, rewards, dones, p_reward=self.p_reward, p_done=self.p_done, seed=self._seed,
                           counter=self._counter, offset=self._host_offset, episode_sums=episode_sums, sums_log=sums_log)
        self._host_offset += 1

    def step_all(self, actions):
        """-> (new_states [n,84,84,C] uint8, rewards [n] fp32, dones [n] fp32), all on the device."""
        torch = self.torch
        if self.native_step:
            new_states, rewards = torch.empty_like(self.states), torch.empty(self.n, dtype=torch.float32, device=self.device)
            dones = torch.empty_like(rewards)
            self.states = torch.empty_like(new_states)             # a fresh tensor: callers may still hold the previous states
            self.step_into(new_states, rewards, dones)
            return new_states, rewards, dones
        u = torch.rand((3, self.n), device=self.device, generator=self._gen)
        rewards = torch.where(u[0] < self.p_reward, torch.where(u[1] < 0.5, 1.0, -1.0), 0.0)
        dones = (u[2] < self.p_done).float()
        new_states = self._frames()
        after_reset = self._frames()
        self.states = torch.where(dones.bool().view(-1, 1, 1, 1), after_reset, new_states)
        return new_states, rewards, dones

    def close(self):
        pass


class BatchedCartPole:
    """`n` cart-poles integrated together on the device (id `CartPoleDevice-v1`): the dynamics, limits and time limit of
    `CartPole`, one tensor expression per step instead of `n` Python calls, states / rewards / dones handed to the agent as
    device tensors (float64 state inside, float32 observations out, like gym's).  Same batch protocol as
    `BatchedSyntheticAtari`; config C1 (PPO on CartPole, n_envs=16, n_steps=128) then runs without a host round trip in
    the rollout."""
    batched = True

    def __init__(self, n, env_id='CartPoleDevice-v1', seed=None, device='cuda:0'):
        import torch
        self.torch, self.n, self.device = torch, int(n), torch.device(device)
        self.spec = _Spec(env_id)
        c = CartPole
        high = np.array([2 * c.X_LIMIT, np.finfo(np.float32).max, 2 * c.THETA_LIMIT, np.finfo(np.float32).max], np.float32)
        self.observation_space = Box(-high, high, (4,), np.float32)
        self.action_space = Discrete(2)
        self.state = self.t = self.states = None
        self.seed(seed)

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return self

    def __iter__(self):
        raise TypeError('a batched environment is not a list of environments: call reset_all() / step_all(actions)')

    def seed(self, seed=None):
        self._gen = self.torch.Generator(device=self.device)
        self._gen.manual_seed(int(seed) if seed is not None else int(np.random.SeedSequence().entropy % (2 ** 63)))
        return [seed]

    def _initial(self, count):
        torch = self.torch
        return torch.rand((count, 4), dtype=torch.float64, device=self.device, generator=self._gen) * 0.1 - 0.05

    def reset_all(self):
        self.state = self._initial(self.n)
        self.t = self.torch.zeros(self.n, dtype=self.torch.int64, device=self.device)
        self.states = self.state.float()
        return self.states

    GRAPH_STATE = ('state', 't', 'states')       # see BatchedSyntheticAtari.GRAPH_STATE

    def rng_state(self):
        return {'gen': self._gen.get_state()}

    def set_rng_state(self, state):
        self._gen.set_state(state['gen'])

    def sync_rng(self):
        pass

    def register_graph(self, graph):
        graph.register_generator_state(self._gen)

    def step_all(self, actions):
        """-> (new_states [n,4] fp32, rewards [n] fp32 (all ones), dones [n] fp32) on the device."""
        torch, c = self.torch, CartPole
        assert self.state is not None, 'Cannot call step_all() before calling reset_all()'
        x, x_dot, th, th_dot = self.state.unbind(1)
        force = torch.where(actions.to(self.device).reshape(-1) >= 0.5, c.FORCE, -c.FORCE).to(torch.float64)
        cos, sin = torch.cos(th), torch.sin(th)
        total, pml = c.CART_MASS + c.POLE_MASS, c.POLE_MASS * c.HALF_LENGTH
        temp = (force + pml * th_dot * th_dot * sin) / total
        th_acc = (c.GRAVITY * sin - cos * temp) / (c.HALF_LENGTH * (4.0 / 3.0 - c.POLE_MASS * cos * cos / total))
        x_acc = temp - pml * th_acc * cos / total
        state = torch.stack([x + c.TAU * x_dot, x_dot + c.TAU * x_acc, th + c.TAU * th_dot, th_dot + c.TAU * th_acc], 1)
        self.t = self.t + 1
        done = (state[:, 0].abs() > c.X_LIMIT) | (state[:, 2].abs() > c.THETA_LIMIT) | (self.t >= c.MAX_STEPS)
        new_states = state.float()
        self.state = torch.where(done.view(-1, 1), self._initial(self.n), state)      # finished episodes start over
        self.t = torch.where(done, torch.zeros_like(self.t), self.t)
        self.states = self.state.float()
        return new_states, torch.ones(self.n, dtype=torch.float32, device=self.device), done.float()

    def close(self):
        pass


class _Gymnasium4Tuple:
    """gymnasium (reset -> (obs, info); step -> 5-tuple) behind the 4-tuple API of the gym version the reference pins."""

    def __init__(self, env):
        self.env = env
        self.observation_space, self.action_space, self.spec = env.observation_space, env.action_space, env.spec

    def reset(self, **kwargs):
        out = self.env.reset(**kwargs)
        return out[0] if isinstance(out, tuple) and len(out) == 2 else out

    def step(self, action):
        out = self.env.step(action)
        if len(out) == 5:
            state, reward, terminated, truncated, info = out
            return state, reward, bool(terminated or truncated), info
        return out

    def seed(self, seed=None):
        if hasattr(self.env, 'seed'):
            return self.env.seed(seed)
        self.env.reset(seed=seed)
        return [seed]

    def close(self):
        self.env.close()


BATCHED = {'SyntheticAtariDevice-v0': BatchedSyntheticAtari, 'CartPoleDevice-v1': BatchedCartPole}

BUILTIN = {
    'CartPole-v1': CartPole,
    'CartPole-v0': CartPole,
    'SyntheticAtari-v0': SyntheticAtari,
    'SyntheticPong-v0': SyntheticAtari,
}


def make(env_name):
    if env_name in BUILTIN:
        return BUILTIN[env_name](env_name)
    for package in ('gym', 'gymnasium'):
        try:
            module = __import__(package)
        except ImportError:
            continue
        return _Gymnasium4Tuple(module.make(env_name))
    if env_name.endswith('NoFrameskip-v4') and os.environ.get('XAGENTS_B200_SYNTHETIC_ATARI') == '1':
        return SyntheticAtari(env_name)
    raise ImportError(f'Cannot create `{env_name}`: neither gym nor gymnasium is installed and it is not one of the '
                      f'built-in environments {sorted(BUILTIN)}')


def create_envs(env_name, n=1, preprocess=True, *args, **kwargs):
    """`n` environments of one id (common.py:145-167).  `preprocess` asks for the reference's AtariWrapper (grayscale +
    84x84 resize + frame skip over raw ALE frames): the built-in synthetic frames are already in the processed
    shape, so it is a no-op for them and an assertion for non-image environments, as in the reference."""
    if env_name in BATCHED:                                        # one device-resident object for all n environments
        made = BATCHED[env_name](n, env_name, device=kwargs.get('device') or 'cuda:0')
        if preprocess:
            shape = made.observation_space.shape
            assert len(shape) == 3, (f'Cannot use AtariWrapper or --preprocess for non-atari environment '
                                     f'{made.spec.id}, with input shape {shape}')
        return made
    envs = [make(env_name) for _ in range(n)]
    if preprocess:
        shape = envs[0].observation_space.shape
        assert len(shape) == 3, (f'Cannot use AtariWrapper or --preprocess for non-atari environment '
                                 f'{envs[0].spec.id}, with input shape {shape}')
        if not isinstance(envs[0], SyntheticAtari):
            raise NotImplementedError('--preprocess of raw emulator frames (cv2 grayscale/resize, frame skipping) happens '
                                      'before the hot path and is not part of this package; wrap the environments yourself')
    return envs
