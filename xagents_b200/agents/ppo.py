"""PPO with the reference's surface (xagents/ppo/agent.py:6-225), hot path on the device.

train_step() = get_batch() (rollout into time-major device buffers -> GAE kernel -> env-major *views*)
followed by run_ppo_epochs().  run_ppo_epochs() drives the prepared pipeline `hotpath.PPOHotPath` IN PLACE over the
agent's own rollout buffers -- the pipeline bench.py measures: one advantage-moments launch for all K*M minibatches,
the permute-gathers back to back on a data stream (double-buffered staging, several minibatches per launch), and per
minibatch [model forward -> fused loss forward+backward reading the rollout scalars through the permutation ->
model backward -> collective C1 -> fused clip+Adam] on the caller's stream, so the HBM-bound byte movement runs under
the network instead of in front of it.

Every replacement point of SURVEY.md §8b keeps its name and argument order.  A subclass (or instance) that overrides
`get_mini_batches`, `update_gradients` or `_normalized_advantages`, or a caller that hands `run_ppo_epochs` arrays
other than the agent's own rollout views, gets the reference's literal loop (`_run_ppo_epochs_generic`: per minibatch
gather -> normalise -> update_gradients) on the same kernels.
"""
import torch

from .. import ops
from ..hotpath import PPOHotPath
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
        self._perm_batch = None          # this train step's K permutations (drawn together at epoch 0)
        self.pipeline_options = {}       # PPOHotPath keywords (gather_mode, gather_chunk, staging, ...) for tuning runs
        self._pipeline = None

    # ------------------------------------------------------------------ returns (ppo/agent.py:48-94)
    def calculate_returns(self, rewards, dones, values=None, selected_critic_logits=None, selected_importance=None):
        dev = lambda x: x if isinstance(x, torch.Tensor) else self._to_device(x, torch.float32)
        rewards, dones, values = dev(rewards), dev(dones), dev(values)
        values = values.reshape(rewards.shape)
        out = self.ro_returns if tuple(rewards.shape) == tuple(self.ro_returns.shape) else None
        return ops.gae_returns(rewards, values, self._bootstrap_values(), dones, self.gamma, self.lam, out=out)

    # ------------------------------------------------------------------ minibatches (ppo/agent.py:139-155)
    def _next_permutation(self, epoch):
        if self.permutation_source is not None:
            perm = self.permutation_source(epoch)
            perm = perm if isinstance(perm, torch.Tensor) else torch.as_tensor(perm)
            return perm.to(self.device, dtype=torch.int32)
        # every epoch's shuffle of a train step in one batched draw + one segmented sort (random 62-bit keys, argsort: what
        # torch.randperm does per call with ~11 launches; four calls were 0.3 ms of a 15 ms update phase)
        if epoch == 0 or self._perm_batch is None or epoch >= self._perm_batch.shape[0]:
            keys = torch.randint(0, 2 ** 62, (self.ppo_epochs, self.batch_size), dtype=torch.int64, device=self.device, generator=self._gen)
            self._perm_batch = keys.argsort(dim=1).to(torch.int32)
        return self._perm_batch[epoch]

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

    # ---- the prepared pipeline (hotpath.PPOHotPath) over the agent's own buffers
    _ROLLOUT_BINDINGS = (('obs', 'ro_states'), ('rewards', 'ro_rewards'), ('values', 'ro_values'), ('dones', 'ro_dones'),
                         ('actions', 'ro_actions'), ('log_probs', 'ro_log_probs'), ('returns', 'ro_returns'))

    def hot_path(self):
        """The agent's PPOHotPath, bound to the current `ro_*` tensors and prepared on torch's current stream."""
        bound = {name: getattr(self, attr) for name, attr in self._ROLLOUT_BINDINGS}
        hp = self._pipeline
        if hp is None:
            options = dict(self.pipeline_options)
            # a network whose first layer fetches every frame by its id needs no gathered copy of the minibatch's frames
            fetches = bool(getattr(self.net, 'reads_through_permutation', False)) and self.obs_dtype == torch.uint8 \
                and tuple(self.input_shape) == (84, 84, 4)
            options.setdefault('obs_gather', not fetches)
            hp = self._pipeline = PPOHotPath(
                self.n_steps, self.n_envs, self.input_shape, self.n_actions, ppo_epochs=self.ppo_epochs,
                mini_batches=self.mini_batches, gamma=self.gamma, lam=self.lam, clip_norm=self.clip_norm,
                entropy_coef=self.entropy_coef, value_loss_coef=self.value_loss_coef,
                advantage_epsilon=self.advantage_epsilon, actor_kind=self.actor_kind, device=self.device, comm=self.comm,
                buffers=bound, **options)
        else:
            moved = {name: t for name, t in bound.items() if getattr(hp, name).data_ptr() != t.data_ptr()}
            if moved:                                              # a rollout feed swapped buffers (double-buffered uploads)
                hp.bind(**moved)
        stream = torch.cuda.current_stream(self.device)
        if hp._calls is None or hp.compute_stream != stream:
            hp.prepare(stream)
        return hp

    def _pipeline_applies(self, *batch):
        names = ('get_mini_batches', 'update_gradients', '_normalized_advantages')
        if any(getattr(type(self), n) is not getattr(PPO, n) or n in self.__dict__ for n in names):
            return False
        buffers = (self.ro_states, self.ro_actions, self.ro_returns, self.ro_values, self.ro_log_probs)
        return len(batch) == 5 and all(isinstance(v, EnvMajorView) and v.tensor.data_ptr() == b.data_ptr()
                                       and v.time_major == (self.n_steps, self.n_envs) for v, b in zip(batch, buffers))

    def run_ppo_epochs(self, states, actions, returns, old_values, old_log_probs):
        if not self._pipeline_applies(states, actions, returns, old_values, old_log_probs):
            return self._run_ppo_epochs_generic(states, actions, returns, old_values, old_log_probs)
        hp, net = self.hot_path(), self.net
        for epoch in range(self.ppo_epochs):                       # every epoch's shuffle up front: one moments launch
            perm = self._next_permutation(epoch)
            if perm.data_ptr() != hp.perms[epoch].data_ptr():      # a source may hand out rows of `hot_path().perms` itself
                hp.perms[epoch].copy_(perm)
        forward_into = getattr(net, 'forward_into', None)

        def forward(i):                                            # minibatch i is staged: model forward on it
            n = hp.mb_rows[i]
            if not hp.obs_gather:                                  # ... or read through the permutation by the first layer
                forward_into(self.ro_states.view((hp.N,) + tuple(self.input_shape)), hp.actor_out[i, :n], hp.critic_out[i, :n], i,
                             idx=hp.minibatch_ids(i), time_major=(self.n_steps, self.n_envs))
                return
            states_mb = hp.mb_obs[hp._mb_place[i][1], hp._mb_place[i][2]:hp._mb_place[i][2] + n]
            if forward_into is not None:
                forward_into(states_mb, hp.actor_out[i, :n], hp.critic_out[i, :n], i)
                return
            actor_out, critic = net.forward(states_mb, training=True)
            hp.actor_out[i, :n].copy_(actor_out.reshape(n, -1))
            hp.critic_out[i, :n].copy_(critic.reshape(n))

        def backward(i):                                           # loss i is queued: back-propagate its output gradients
            n = hp.mb_rows[i]
            net.backward_and_step(hp.d_actor[:n], hp.d_values[:n], self.grad_norm)

        hp.run(gae=False, before_loss=forward, after_loss=backward)
        self.loss_history[:] = [hp.scalars[i] for i in range(hp.n_mb)]

    def _run_ppo_epochs_generic(self, states, actions, returns, old_values, old_log_probs):
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
        if not isinstance(returns, torch.Tensor):
            returns = self._to_device(returns, torch.float32)
        if returns.data_ptr() != self.ro_returns.data_ptr():       # an overriding calculate_returns made its own tensor
            self.ro_returns = returns.contiguous()
        return self.concat_step_batches(states, actions, self.ro_returns, values, log_probs)

    def train_step(self):
        batch = self.get_batch()
        self.run_ppo_epochs(*batch)
