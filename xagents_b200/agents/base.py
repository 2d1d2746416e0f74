"""Driver-side mirror of the reference's BaseAgent / OnPolicy for the on-policy path.

Keeps the constructor keywords, attribute names, `fit()` contract and assertion messages of
`xagents/base.py:22-128, 566-593` so callers and tests written against the reference keep working, but
holds only what the rollout-to-update path needs: environment stepping, episode bookkeeping and the
stop conditions.  Checkpointing, parquet history, wandb and optuna hooks are out of scope
and are rejected loudly rather than silently ignored.
"""
import random
from collections import deque
from time import perf_counter

import numpy as np
import torch


class EnvMajorView:
    """A time-major device field [T, E, ...] standing for the env-major flat array [T*E, ...] that
    BaseAgent.concat_step_batches (xagents/base.py:549-564) would have copied out.  The gathers and
    losses take (tensor, T, E) and remap indices instead, so the copy never happens."""

    def __init__(self, tensor, n_steps, n_envs):
        self.tensor, self.n_steps, self.n_envs = tensor, int(n_steps), int(n_envs)

    @property
    def shape(self):
        return (self.n_steps * self.n_envs,) + tuple(self.tensor.shape[2:])

    def __len__(self):
        return self.n_steps * self.n_envs

    @property
    def time_major(self):
        return (self.n_steps, self.n_envs)

    def materialize(self):
        """The actual env-major copy (for callers that insist on a flat array)."""
        from .. import ops
        ids = torch.arange(len(self), dtype=torch.int32, device=self.tensor.device)
        src = self.tensor if self.tensor.dim() > 2 else self.tensor.reshape(self.n_steps, self.n_envs, 1)
        out = ops.gather_rows(src, ids, time_major=self.time_major)
        return out if self.tensor.dim() > 2 else out.reshape(-1)


class BaseAgent:
    def __init__(self, envs, model, checkpoints=None, reward_buffer_size=100, n_steps=1, gamma=0.99,
                 display_precision=2, seed=None, log_frequency=None, history_checkpoint=None,
                 plateau_reduce_factor=0.9, plateau_reduce_patience=10, early_stop_patience=3,
                 divergence_monitoring_steps=None, quiet=False, trial=None, device='cuda:0'):
        assert envs, 'No environments given'
        for name, value in (('checkpoints', checkpoints), ('history_checkpoint', history_checkpoint), ('trial', trial)):
            if value is not None:
                raise NotImplementedError(f'{name}: checkpoint / history / tuning hooks are outside the hot-path '
                                          f'package (SURVEY.md §2); drive them from the caller')
        # one device-resident object for all environments (envs.BatchedSyntheticAtari): states, rewards and dones are
        # device tensors and a rollout step has no host round trip; a list of gym-style environments otherwise
        self.batched = bool(getattr(envs, 'batched', False))
        self.n_envs = len(envs)
        self.envs = envs
        self.model = model
        self.checkpoints = checkpoints
        self.total_rewards = deque(maxlen=reward_buffer_size)
        self.n_steps = n_steps
        self.gamma = gamma
        self.display_precision = display_precision
        self.seed = seed
        self.output_models = [self.model]
        self.log_frequency = log_frequency or self.n_envs
        self.history_checkpoint = history_checkpoint
        self.plateau_reduce_factor = plateau_reduce_factor
        self.plateau_reduce_patience = plateau_reduce_patience
        self.early_stop_patience = early_stop_patience
        self.divergence_monitoring_steps = divergence_monitoring_steps
        self.quiet = quiet
        self.trial = trial
        self.device = torch.device(device)
        self.target_reward = None
        self.max_steps = None
        self.input_shape = tuple(self.envs[0].observation_space.shape)
        self.best_reward = -float('inf')
        self.mean_reward = -float('inf')
        self.states = [np.array(0)] * self.n_envs
        self.dones = [False] * self.n_envs
        if self.batched:
            self.dones = torch.zeros(self.n_envs, dtype=torch.float32, device=self.device)
            self._episode_sums = torch.zeros(self.n_envs, dtype=torch.float32, device=self.device)
            self._episode_log = []       # per step: (dones, episode sums at that step), read back once per train step
        self.steps = 0
        self.frame_speed = 0
        self.last_reset_step = 0
        self.training_start_time = None
        self.last_reset_time = None
        self.games = 0
        self.episode_rewards = np.zeros(self.n_envs)
        self.done_envs = 0
        self.plateau_count = 0
        self.early_stop_count = 0
        if seed:
            self.set_seeds(seed)
        self.reset_envs()
        self.set_action_count()
        self.img_inputs = len(tuple(self.states[0].shape) if self.batched else np.shape(self.states[0])) >= 2   # base.py:120
        self.display_titles = ('time', 'steps', 'games', 'speed', 'mean reward', 'best reward')

    # ------------------------------------------------------------------ small helpers
    def display_message(self, *args, **kwargs):
        if not self.quiet:
            print(*args, **kwargs)

    def set_seeds(self, seed):
        torch.manual_seed(seed)
        np.random.seed(seed)
        random.seed(seed)
        if self.batched:
            self.envs.seed(seed)
            return
        for env in self.envs:
            if hasattr(env, 'seed'):
                env.seed(seed)
            if hasattr(env.action_space, 'seed'):
                env.action_space.seed(seed)

    def reset_envs(self):
        if self.batched:
            self.states = self.envs.reset_all()
            return
        for i, env in enumerate(self.envs):
            self.states[i] = env.reset()

    def set_action_count(self):
        space = self.envs[0].action_space
        self.discrete = hasattr(space, 'n')
        if self.discrete:
            self.n_actions = int(space.n)
        else:
            assert hasattr(space, 'shape') and len(space.shape) >= 1, f'Expected a Box or Discrete action space, got {space}'
            self.n_actions = int(space.shape[0])

    def get_states(self):
        return self.states if self.batched else np.array(self.states)

    def get_dones(self):
        return self.dones if self.batched else np.array(self.dones, np.float32)

    def _step_batched(self, actions, get_observation):
        """step_envs for a device-resident batched environment: same contract, device tensors, no per-env loop.  Episode
        bookkeeping (total_rewards, games, done_envs) is deferred: the per-step done flags and episode sums stay on the
        device and are read back once per train step by `_flush_episode_log`."""
        previous = self.states
        new_states, rewards, dones = self.envs.step_all(actions)
        self.states, self.dones = self.envs.states, dones
        self._episode_sums += rewards
        self._episode_log.append((dones, self._episode_sums.clone()))
        self._episode_sums *= 1.0 - dones
        self.steps += self.n_envs
        return [previous, actions, rewards, dones, new_states] if get_observation else []

    def _flush_episode_log(self):
        if not self.batched or not self._episode_log:
            return
        dones = torch.stack([d for d, _ in self._episode_log]).cpu().numpy()
        sums = torch.stack([s for _, s in self._episode_log]).cpu().numpy()
        self._episode_log.clear()
        for step_dones, step_sums in zip(dones, sums):             # in step order, envs in index order: the loop's order
            for total in step_sums[step_dones > 0]:
                self.done_envs += 1
                self.total_rewards.append(float(total))
                self.games += 1

    # ------------------------------------------------------------------ environments
    def step_envs(self, actions, get_observation=False, store_in_buffers=False):
        """Step every environment once (xagents/base.py:388-426 semantics: the returned new_state of a
        finished episode is the terminal one while `self.states[i]` already holds the reset state)."""
        assert not store_in_buffers, 'replay buffers belong to the off-policy agents (out of scope)'
        if self.batched:
            return self._step_batched(actions, get_observation)
        columns = [[] for _ in range(5)]
        for i, (env, action) in enumerate(zip(self.envs, actions)):
            previous = self.states[i]
            new_state, reward, done, _ = env.step(action)
            self.states[i], self.dones[i] = new_state, done
            self.episode_rewards[i] += reward
            if get_observation:
                for col, item in zip(columns, (previous, action, reward, done, new_state)):
                    col.append(item)
            if done:
                self.done_envs += 1
                self.total_rewards.append(self.episode_rewards[i])
                self.games += 1
                self.episode_rewards[i] = 0
                self.states[i] = env.reset()
            self.steps += 1
        if not get_observation:
            return []
        # The reference casts every column to float32 (base.py:426), frames included: 4x the bytes for the same integers.
        # uint8 frames stay uint8 here (as they do in the rollout buffers); at 256 Atari-shaped environments that is
        # 0.9 ms instead of 13.5 ms of host work per step, and a quarter of the H2D bytes.
        keep = [np.asarray(col[0]).dtype == np.uint8 for col in (columns[0], columns[4])]
        dtypes = [np.uint8 if keep[0] else np.float32, np.float32, np.float32, np.float32, np.uint8 if keep[1] else np.float32]
        return [np.array(col, dtype) for col, dtype in zip(columns, dtypes)]

    @staticmethod
    def concat_step_batches(*args):
        """[T, E, ...] -> env-major [T*E, ...] for each argument (xagents/base.py:549-564).  Device tensors
        come back as `EnvMajorView`s (no copy); host arrays are copied like the reference does."""
        out = []
        for arg in args:
            if isinstance(arg, torch.Tensor) and arg.is_cuda:
                t = arg if arg.dim() > 1 else arg.unsqueeze(-1)
                out.append(EnvMajorView(t, t.shape[0], t.shape[1]))
            else:
                a = np.asarray(arg)
                a = a[:, None] if a.ndim == 1 else a
                out.append(np.ascontiguousarray(np.swapaxes(a, 0, 1)).reshape((-1,) + a.shape[2:]))
        return out

    # ------------------------------------------------------------------ driver
    def update_metrics(self):
        """base.py:213-230, 260-291: best reward follows the mean (the reference's `checkpoint()`; saving weights is out of
        scope), plateau / early-stop counters under `divergence_monitoring_steps`, learning-rate reduction on a plateau, frame
        speed, then the new mean reward -- in the reference's order, so the plateau test sees the PREVIOUS mean."""
        self._flush_episode_log()
        if self.mean_reward > self.best_reward:
            self.plateau_count = 0
            self.early_stop_count = 0
            self.display_message(f'Best reward updated: {self.best_reward} -> {self.mean_reward}')
        self.best_reward = max(self.mean_reward, self.best_reward)
        if (self.divergence_monitoring_steps and self.steps >= self.divergence_monitoring_steps
                and self.mean_reward <= self.best_reward):
            self.plateau_count += 1
        if self.plateau_count >= self.plateau_reduce_patience:
            for model in self.output_models:                       # adapters carry the optimiser's learning rate as `.lr`
                net = getattr(self, 'net', None) if not hasattr(model, 'lr') else model
                if net is not None and hasattr(net, 'lr'):
                    new_lr = net.lr * self.plateau_reduce_factor
                    self.display_message(f'Learning rate reduced {net.lr} -> {new_lr}')
                    net.lr = new_lr
            self.plateau_count = 0
            self.early_stop_count += 1
        now = perf_counter()
        since = self.last_reset_time if self.last_reset_time is not None else now        # metrics asked for outside fit()
        self.frame_speed = (self.steps - self.last_reset_step) / max(now - since, 1e-9)
        self.last_reset_step, self.last_reset_time = self.steps, now
        self.mean_reward = (float(np.around(np.mean(self.total_rewards), self.display_precision)) if self.total_rewards
                            else -float('inf'))

    def display_metrics(self):
        elapsed = perf_counter() - self.training_start_time
        values = (f'{elapsed:.0f}s', self.steps, self.games, f'{self.frame_speed:.0f} steps/s',
                  round(self.mean_reward, self.display_precision), round(self.best_reward, self.display_precision))
        self.display_message(', '.join(f'{t}: {v}' for t, v in zip(self.display_titles, values)))

    def check_episodes(self):
        self._flush_episode_log()
        if self.done_envs >= self.log_frequency:
            self.update_metrics()
            self.display_metrics()
            self.done_envs = 0

    def training_done(self):
        self._flush_episode_log()
        mean_reward, steps = self.mean_reward, self.steps
        comm = getattr(self, 'comm', None)
        if comm is not None and comm.world_size > 1:
            # sharded environments: every rank must take the same decision at the same train step (the updates contain
            # collectives), so the stop conditions read the job-wide episode mean and the job-wide step count
            total, count, steps = comm.sum_over_ranks([float(np.sum(self.total_rewards)), len(self.total_rewards), self.steps])
            mean_reward = total / count if count else -float('inf')
            steps = int(steps)
        if self.early_stop_count >= self.early_stop_patience:      # base.py:333-335
            self.display_message('Early stopping')
            return True
        if self.target_reward is not None and mean_reward >= self.target_reward:
            self.display_message(f'Reward achieved in {steps} steps')
            return True
        if self.max_steps and steps >= self.max_steps:
            self.display_message(f'Maximum steps exceeded')
            return True
        return False

    def at_step_start(self):
        pass

    def at_step_end(self):
        pass

    def train_step(self):
        raise NotImplementedError(f'train_step() should be implemented by {self.__class__.__name__} subclasses')

    def fit(self, target_reward=None, max_steps=None, monitor_session=None):
        """Same contract as xagents/base.py:566-593: one of target_reward / max_steps is required;
        loops at_step_start(); train_step(); at_step_end() until a stop condition holds."""
        assert target_reward or max_steps, '`target_reward` or `max_steps` should be specified when fit() is called'
        if monitor_session is not None:
            raise NotImplementedError('monitor_session: wandb monitoring is outside the hot-path package')
        self.target_reward, self.max_steps = target_reward, max_steps
        self.training_start_time = self.last_reset_time = perf_counter()
        while True:
            self.check_episodes()
            if self.training_done():
                break
            self.at_step_start()
            self.train_step()
            self.at_step_end()


class OnPolicy(BaseAgent):
    """Marker base of the on-policy agents (xagents/base.py:656-670)."""
