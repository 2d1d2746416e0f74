"""Rollouts that are produced outside the agent, fed to `A2C/PPO.train_step` from HOST memory.

The reference's `get_batch` (xagents/a2c/agent.py:96-139) steps Python environments and grows host lists that
`np.asarray(..., np.float32)` then copies (ppo/agent.py:202-210).  When the rollout already exists on the host --
actors in other processes, recorded data, the benchmark's synthetic rollouts -- `HostRolloutFeed` is the agent's
`rollout_source`: it uploads rollout k+1 from pinned host buffers on a copy stream into a second set of device
buffers WHILE train step k runs, then points the agent's `ro_*` attributes at the uploaded set (the prepared
pipeline re-binds in place, no device copy).  uint8 frames travel as uint8 (a quarter of the reference's fp32 bytes).

    feed = HostRolloutFeed(agent, source)      # source(k) -> dict of pinned host tensors, or None when exhausted
    agent.rollout_source = feed
    agent.train_step()                         # H2D of rollout k (already in flight) + GAE + epochs
"""
import torch


class HostRolloutFeed:
    FIELDS = (('obs', 'ro_states'), ('rewards', 'ro_rewards'), ('values', 'ro_values'), ('dones', 'ro_dones'),
              ('actions', 'ro_actions'), ('log_probs', 'ro_log_probs'))

    def __init__(self, agent, source, extras=None, slots=2, on_switch=None, chunks=4):
        """`source(k)`: dict with time-major host tensors 'obs', 'rewards', 'values', 'dones' [T+1, E], 'actions',
        'log_probs', 'last_values' [E] (+ every key of `extras`).  `extras`: {name: (shape, dtype)} additional device
        buffers uploaded per rollout (precomputed model outputs, permutations); the current set is `feed.current`.
        `on_switch(current)` runs after the agent was pointed at a new set.  `chunks`: the observation upload is cut
        into this many copies so that it interleaves with the small fields of the same rollout on the copy engine."""
        self.agent, self.source, self.on_switch, self.chunks = agent, source, on_switch, max(1, int(chunks))
        dev = agent.device
        self.device = dev
        self.sets = []
        for s in range(slots):
            bufs = {}
            for name, attr in self.FIELDS:
                own = getattr(agent, attr)
                bufs[name] = own if s == 0 else torch.empty_like(own)          # set 0 = the agent's own buffers
            bufs['last_values'] = torch.empty((agent.n_envs,), dtype=torch.float32, device=dev)
            for name, (shape, dtype) in (extras or {}).items():
                bufs[name] = torch.empty(shape, dtype=dtype, device=dev)
            self.sets.append(bufs)
        self.copy_stream = torch.cuda.Stream(dev)
        self.uploaded = [torch.cuda.Event() for _ in range(slots)]
        self.released = [torch.cuda.Event() for _ in range(slots)]
        self.in_flight = [None] * slots          # rollout number whose upload was issued into the set
        self.k = 0
        self.current = None
        self.h2d_bytes_per_rollout = sum(t.numel() * t.element_size() for t in self.sets[0].values())

    def _upload(self, k, s):
        host = self.source(k)
        if host is None:
            return False
        dst = self.sets[s]
        missing = [name for name in dst if name not in host]
        assert not missing, f'rollout source did not supply {missing}'
        cs = self.copy_stream
        cs.wait_event(self.released[s])                     # the train step that read this set has finished
        with torch.cuda.stream(cs):
            for name, t in dst.items():
                src = torch.as_tensor(host[name])
                assert tuple(src.shape) == tuple(t.shape) and src.dtype == t.dtype, (
                    f'{name}: host {tuple(src.shape)} {src.dtype} vs device {tuple(t.shape)} {t.dtype}')
                if name == 'obs' and self.chunks > 1 and t.shape[0] >= self.chunks:
                    for part_dst, part_src in zip(t.chunk(self.chunks), src.chunk(self.chunks)):
                        part_dst.copy_(part_src, non_blocking=True)
                else:
                    t.copy_(src, non_blocking=True)
            self.uploaded[s].record(cs)
        self.in_flight[s] = k
        return True

    def __call__(self, agent):
        assert agent is self.agent
        n, k = len(self.sets), self.k
        s = k % n
        stream = torch.cuda.current_stream(self.device)
        # everything queued so far (the previous train step) has stopped reading the OTHER sets once this event fires
        for other in range(n):
            if other != s:
                self.released[other].record(stream)
        if self.in_flight[s] != k:
            self.released[s].record(stream)
            assert self._upload(k, s), f'rollout source is exhausted at rollout {k}'
        if n > 1:
            self._upload(k + 1, (k + 1) % n)                 # travels under this train step's kernels
        stream.wait_event(self.uploaded[s])
        cur = self.current = self.sets[s]
        for name, attr in self.FIELDS:
            setattr(agent, attr, cur[name])
        if self.on_switch is not None:
            self.on_switch(cur)
        self.k += 1
        return cur['last_values']


class FedEnvs:
    """The environment side of an agent whose rollouts come from a feed: spaces and a count, nothing to step.  Looks like
    the batched device environments (`envs.BatchedSyntheticAtari`) to BaseAgent."""
    batched = True

    def __init__(self, n, obs_shape, obs_dtype, n_actions, device='cuda:0', env_id='FedRollouts-v0'):
        import numpy as np

        from . import envs
        self.n, self.device = int(n), torch.device(device)
        self.spec = envs._Spec(env_id)
        np_dtype = np.uint8 if obs_dtype in (torch.uint8, 'uint8', np.uint8) else np.float32
        self.observation_space = envs.Box(0, 255, tuple(obs_shape), np_dtype)
        self.action_space = envs.Discrete(n_actions)
        self._dtype = torch.uint8 if np_dtype == np.uint8 else torch.float32
        self.states = None

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return self

    def seed(self, seed=None):
        return [seed]

    def reset_all(self):
        self.states = torch.zeros((self.n,) + self.observation_space.shape, dtype=self._dtype, device=self.device)
        return self.states

    def step_all(self, actions):
        raise RuntimeError('FedEnvs cannot be stepped: set agent.rollout_source (feeds.HostRolloutFeed)')

    def close(self):
        pass
