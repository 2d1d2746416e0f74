"""PPO with the reference's surface (xagents/ppo/agent.py:6-225), hot path on the device.

train_step() = get_batch() (rollout into time-major device buffers -> GAE kernel -> env-major *views*)
followed by run_ppo_epochs() (per minibatch: permute-gather kernel, advantage moments + normalisation,
model forward, fused loss forward+backward, model backward, fused clip+Adam).  Every replacement point of
SURVEY.md §8b keeps its name and argument order, so subclasses that override one of them still compose.
"""
import torch

from .. import ops
from .a2c import A2C
from .base import EnvMajorView


class PPO(A2C):
    def __init__(self, envs, model, lam=0.95, ppo_epochs=4, mini_batches=4, advantage_epsilon=1e-8, clip_norm=0.1,
                 **kwargs):
        super().__init__(envs, model, **kwargs)
        self.lam = lam
        self.ppo_epochs = ppo_epochs
        self.mini_batches = mini_batches
        self.advantage_epsilon = advantage_epsilon
        self.clip_norm = clip_norm
        self.batch_size = self.n_envs * self.n_steps
        self.mini_batch_size = self.batch_size // self.mini_batches
        assert self.mini_batch_size > 0, (
            f'Invalid batch size to mini-batch size ratio {self.batch_size}: {self.mini_batches}')
        self.permutation_source = None   # tests / parity runs: callable(epoch) -> int32 permutation of range(N)
        self.loss_history = []           # device tensors [4] = loss, pg, value loss, entropy per update
        self._workspace = ops.loss_workspace(self.mini_batch_size + self.mini_batches, self.device)

    # ------------------------------------------------------------------ returns (ppo/agent.py:48-94)
    def calculate_returns(self, rewards, dones, values=None, selected_critic_logits=None, selected_importance=None):
        dev = lambda x: x if isinstance(x, torch.Tensor) else self._to_device(x, torch.float32)
        rewards, dones, values = dev(rewards), dev(dones), dev(values)
        values = values.reshape(rewards.shape)
        return ops.gae_returns(rewards, values, self._bootstrap_values(), dones, self.gamma, self.lam)

    # ------------------------------------------------------------------ minibatches (ppo/agent.py:139-155)
    def _next_permutation(self, epoch):
        if self.permutation_source is not None:
            perm = self.permutation_source(epoch)
            perm = perm if isinstance(perm, torch.Tensor) else torch.as_tensor(perm)
            return perm.to(self.device, dtype=torch.int32)
        return torch.randperm(self.batch_size, device=self.device, generator=self._gen).to(torch.int32)

    def get_mini_batches(self, *args):
        """Yield, for every epoch and every `range(0, N, B)` slice (trailing short one included), the list of
        gathered items.  Lazily: `list(...)` reproduces the reference's up-front list, iterating keeps one
        minibatch resident.  Items may be `EnvMajorView`s (time-major buffers, remapped) or flat device tensors."""
        views = list(args)
        for epoch in range(self.ppo_epochs):
            perm = self._next_permutation(epoch)
            for lo in range(0, self.batch_size, self.mini_batch_size):
                idx = perm[lo:lo + self.mini_batch_size]
                yield [self._gather(item, idx) for item in views]

    @staticmethod
    def _gather(item, idx):
        if isinstance(item, EnvMajorView):
            src = item.tensor if item.tensor.dim() > 2 else item.tensor.reshape(item.n_steps, item.n_envs, 1)
            out = ops.gather_rows(src, idx, time_major=item.time_major)
            return out if item.tensor.dim() > 2 else out.reshape(-1)
        src = item if item.dim() > 1 else item.reshape(-1, 1)
        out = ops.gather_rows(src, idx)
        return out if item.dim() > 1 else out.reshape(-1)

    # ------------------------------------------------------------------ updates (ppo/agent.py:96-137, 157-191)
    def update_gradients(self, states, actions, old_values, returns, old_log_probs, advantages):
        actor_out, critic = self.net.forward(states, training=True)
        scalars, d_actor, d_values, _ = ops.ppo_loss(
            actor_out, critic, actions, old_log_probs, old_values, returns, advantages=advantages,
            clip_norm=self.clip_norm, entropy_coef=self.entropy_coef, value_loss_coef=self.value_loss_coef,
            advantage_epsilon=self.advantage_epsilon, actor_kind=self.actor_kind, workspace=self._workspace)
        self.loss_history.append(scalars)
        self.net.backward_and_step(d_actor, d_values, self.grad_norm)

    def _normalized_advantages(self, returns_mb, old_values_mb):
        n = returns_mb.shape[0]
        moments = ops.adv_moments(returns_mb, old_values_mb, None, [0, n])
        if self.comm is not None and self.comm.world_size > 1:        # collective C2: global minibatch statistics
            parts = torch.empty((self.comm.world_size,) + tuple(moments.shape), dtype=moments.dtype, device=moments.device)
            self.comm.all_gather_moments(parts, moments)
            moments = parts.reshape(self.comm.world_size, -1)
        else:
            moments = moments[0]
        return ops.normalize_advantages(returns_mb, old_values_mb, self.advantage_epsilon, moments=moments)

    def run_ppo_epochs(self, states, actions, returns, old_values, old_log_probs):
        self.loss_history.clear()
        for states_mb, actions_mb, returns_mb, old_values_mb, old_log_probs_mb in self.get_mini_batches(
                states, actions, returns, old_values, old_log_probs):
            advantages_mb = self._normalized_advantages(returns_mb, old_values_mb)
            self.update_gradients(states_mb, actions_mb, old_values_mb, returns_mb, old_log_probs_mb, advantages_mb)

    # ------------------------------------------------------------------ batch + step (ppo/agent.py:193-225)
    def get_batch(self):
        """[states, actions, returns, values, log probs], env-major, as views over the time-major buffers."""
        states, rewards, actions, values, dones, log_probs, *_ = super().get_batch()
        returns = self.calculate_returns(rewards, dones, values)
        self.ro_returns = returns
        return self.concat_step_batches(states, actions, returns, values, log_probs)

    def train_step(self):
        batch = self.get_batch()
        self.run_ppo_epochs(*batch)
