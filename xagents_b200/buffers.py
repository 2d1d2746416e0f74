"""Replay storage for ACER (xagents/utils/buffers.py:8-99, xagents/acer/agent.py:128-170).

The reference keeps one `ReplayBuffer1` per environment: a host deque of trajectory tuples (T+1 frames as LazyFrames,
T rewards / actions / dones, T x A behaviour probabilities); a replay batch is one random trajectory per environment,
concatenated on the host and shipped to the device again.  Here the trajectories never leave HBM:

* `ReplayBuffer1` keeps the reference's constructor, assertions and bookkeeping attributes (`size`, `initial_size`,
  `batch_size`, `current_size`) -- it is the object `create_buffers` hands to `ACER(envs, model, buffers)`;
* `DeviceTrajectoryRing` is where the data lives: time-major slots `[S, T+1, E, ...]` (uint8 frames stay uint8)
  that the rollout writes in place, and a replay batch is ONE launch of the hot path's row-gather kernel per field with
  the row ids `((slot_e * rows + t) * E + e)` -- each environment draws its own slot, like one `get_sample()` per buffer.
  625 slots x 21 frames x 16 envs of 84x84x4 (the CLI defaults) are 5.9 GB of the 180 GB.
"""
import numpy as np
import torch

from . import ops


class BaseBuffer:
    def __init__(self, size, initial_size=None, batch_size=32):
        assert initial_size is None or initial_size > 0, f'Buffer initial size should be > 0, got {initial_size}'
        assert size > 0, f'Buffer size should be > 0,  got {size}'
        assert batch_size > 0, f'Buffer batch size should be > 0, got {batch_size}'
        assert batch_size <= size, f'Buffer batch size `{batch_size}` should be <= size `{size}`'
        if initial_size:
            assert size >= initial_size, 'Buffer initial size exceeds max size'
        self.size = size
        self.initial_size = initial_size or size
        self.batch_size = batch_size
        self.current_size = 0

    def append(self, *args):
        raise NotImplementedError(f'append() should be implemented by {self.__class__.__name__} subclasses')

    def get_sample(self):
        raise NotImplementedError(f'get_sample() should be implemented by {self.__class__.__name__} subclasses')


class ReplayBuffer1(BaseBuffer):
    """Per-environment handle on the device ring: same sizes and counters as the reference's deque buffer; the
    trajectories themselves are rows of `DeviceTrajectoryRing` (append / get_sample go through the agent)."""

    def __init__(self, size, **kwargs):
        super().__init__(size, **kwargs)
        self.ring = None                 # set by the agent that owns the storage
        self.env_index = None

    def append(self, *args):
        raise NotImplementedError('trajectories are written into the device ring by the rollout itself (ACER.get_batch)')

    def get_sample(self):
        assert self.ring is not None, 'this buffer is not attached to an agent'
        return self.ring.sample_slots()[self.env_index]


def create_buffers(agent_id, max_size, batch_size, n_envs, initial_size=None, as_total=True):
    """common.py:497-565 for the agent this package mirrors with a buffer (ACER: batch size 1)."""
    assert agent_id == 'acer', f'replay buffers of `{agent_id}` belong to the off-policy agents (outside the hot path)'
    initial_size = initial_size or max_size
    if as_total:
        max_size //= n_envs
        initial_size //= n_envs
        batch_size //= n_envs
    return [ReplayBuffer1(max_size, initial_size=initial_size, batch_size=1) for _ in range(n_envs)]


class DeviceTrajectoryRing:
    def __init__(self, slots, n_steps, n_envs, obs_shape, obs_dtype, n_actions, device, seed=None):
        S, T, E, dev, f32 = int(slots), int(n_steps), int(n_envs), device, torch.float32
        assert S * (T + 1) * E < 2 ** 31, 'ring rows must be addressable with int32 ids'
        self.slots, self.n_steps, self.n_envs = S, T, E
        self.states = torch.empty((S, T + 1, E) + tuple(obs_shape), dtype=obs_dtype, device=dev)
        self.rewards = torch.empty((S, T, E), dtype=f32, device=dev)
        self.actions = torch.empty((S, T, E), dtype=f32, device=dev)
        self.dones = torch.empty((S, T, E), dtype=f32, device=dev)
        self.probs = torch.empty((S, T, E, n_actions), dtype=f32, device=dev)
        self.written = 0                 # trajectories stored so far (per environment)
        self._rng = np.random.default_rng(seed)
        self._t = torch.arange(T + 1, dtype=torch.int64, device=dev).view(-1, 1)
        self._e = torch.arange(E, dtype=torch.int64, device=dev).view(1, -1)

    @property
    def current_size(self):
        return min(self.written, self.slots)

    def write_slot(self):
        """Index of the slot the next rollout fills (oldest trajectory is overwritten: deque(maxlen=size))."""
        return self.written % self.slots

    def commit(self):
        self.written += 1

    def slot_fields(self, slot):
        return self.states[slot], self.rewards[slot], self.actions[slot], self.dones[slot], self.probs[slot]

    def sample_slots(self):
        """One stored trajectory per environment, uniformly (random.sample(main_buffer, 1) per buffer)."""
        return self._rng.integers(0, self.current_size, self.n_envs)

    def gather(self, slots):
        """Time-major batch [T(+1), E, ...] made of trajectory `slots[e]` for environment e: one gather launch per field."""
        T, E = self.n_steps, self.n_envs
        s = torch.as_tensor(np.asarray(slots, np.int64), device=self.states.device).view(1, -1)
        ids_states = ((s * (T + 1) + self._t) * E + self._e).reshape(-1).to(torch.int32)
        ids_steps = ((s * T + self._t[:T]) * E + self._e).reshape(-1).to(torch.int32)
        obs_shape = tuple(self.states.shape[3:])
        states = ops.gather_rows(self.states.view((-1,) + obs_shape), ids_states).view((T + 1, E) + obs_shape)
        rewards, actions, dones = (ops.gather_rows(f.view(-1, 1), ids_steps).view(T, E) for f in (self.rewards, self.actions, self.dones))
        probs = ops.gather_rows(self.probs.view(-1, self.probs.shape[-1]), ids_steps).view(T, E, -1)
        return states, rewards, actions, dones, probs
