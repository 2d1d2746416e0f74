"""ACER with the reference's surface (xagents/acer/agent.py:9-387) on the hot-path kernels.

Differences in mechanism, not in arithmetic:

* The rollout writes straight into a slot of the device trajectory ring (`buffers.DeviceTrajectoryRing`); there is no
  `store_batch` copy, no LazyFrames, no host concatenation.  A replay batch (one random stored trajectory per
  environment, `concat_buffer_samples`, base.py:428-452) is one row-gather launch per field.
* Batches stay TIME-major ([T(+1), E, ...]) instead of the reference's env-major flat arrays: every ACER reduction is a
  mean over all samples and the only order-dependent step, Retrace, is the reverse-scan kernel `xa_retrace_f32`
  (bit-exact against the reference's own `calculate_returns`, tests/golden/acer_retrace.npz), which is time-major.
* Returns carry no gradient in the reference (they enter the losses under `stop_gradient`, acer/agent.py:229-241), so
  the scan runs outside autograd.  The losses' derivatives w.r.t. the two model outputs (action probabilities and
  per-action critic values) and the trust-region projection of acer/agent.py:264-291 are taken on those small [N, A]
  tensors; the network then back-propagates the two output gradients and the fused clip+Adam kernel applies them.
* The averaged policy network is the flat parameter vector's exponential moving average, applied functionally
  (`tf.train.ExponentialMovingAverage` semantics: the shadow starts as a copy at the first `apply`).  Before the
  first update the reference's `clone_model` holds freshly initialised weights; here it holds a copy of the model's.
"""
import numpy as np
import torch

from .. import ops
from ..buffers import DeviceTrajectoryRing
from .a2c import A2C


class ACER(A2C):
    def __init__(self, envs, model, buffers, ema_alpha=0.99, replay_ratio=4, epsilon=1e-6, importance_c=10.0, delta=1,
                 trust_region=True, **kwargs):
        super().__init__(envs, model, **kwargs)
        assert self.discrete, (f'Invalid environment: {getattr(getattr(envs[0], "spec", None), "id", envs[0])}. ACER supports '
                               f'environments with a Discrete action space only, got {envs[0].action_space}')
        assert len(buffers) == self.n_envs, f'Expected {self.n_envs} buffers, got {len(buffers)}'
        assert buffers[0].batch_size == 1, f'Buffer batch size should be 1 for ACER, got {buffers[0].batch_size}'
        assert hasattr(self.net, 'flat_param'), 'ACER needs a flat-parameter adapter (TorchModel) for its averaged network'
        self.buffers = buffers
        self.ema_alpha = ema_alpha
        self.replay_ratio = replay_ratio
        self.epsilon = epsilon
        self.importance_c = importance_c
        self.delta = delta
        self.trust_region = trust_region
        self.buffer_current_size = 0
        self.batch_dtypes = ['uint8', 'float32', 'int32', 'float32', 'float32']
        T, E, A = self.n_steps, self.n_envs, self.n_actions
        self.batch_shapes = [(E * (T + 1), *self.input_shape), (E * T,), (E * T,), (E * T,), (E * T, A)]
        self.ring = DeviceTrajectoryRing(buffers[0].size, T, E, self.input_shape, self.obs_dtype, A, self.device, seed=self.seed)
        for i, buffer in enumerate(buffers):
            buffer.ring, buffer.env_index = self.ring, i
        self.avg_flat = self.net.flat_param.clone()               # the averaged network's weights
        self._ema_started = False
        self._views = []
        off = 0
        for name, p in self.net.module.named_parameters():
            if p.requires_grad:
                self._views.append((name, off, tuple(p.shape)))
                off += p.numel()
        self._poisson = np.random.default_rng(self.seed)
        self.last_losses = None

    # ------------------------------------------------------------------ rollout into the ring (acer/agent.py:146-170)
    def get_batch(self):
        """Run n_steps in every environment, writing the trajectory into the ring slot it will be replayed from.
        Returns the slot's time-major fields [states (T+1 rows), rewards, actions, dones (after each step), actor output]."""
        slot = self.ring.write_slot()
        states, rewards, actions, dones, probs = self.ring.slot_fields(slot)
        step_states = self.get_states()
        for t in range(self.n_steps):
            states_d = self._to_device(step_states, self.obs_dtype)
            step_actions, _, _, _, actor_out = self.get_model_outputs(states_d, training=False, step=t)
            states[t].copy_(states_d)
            actions[t].copy_(step_actions)
            probs[t].copy_(actor_out)
            *_, step_rewards, step_dones, step_states = self.step_envs(self._env_actions(step_actions), True, False)
            rewards[t].copy_(self._to_device(step_rewards))
            dones[t].copy_(self._to_device(step_dones))
        states[self.n_steps].copy_(self._to_device(self.get_states(), self.obs_dtype))   # acer/agent.py:160
        self.ring.commit()
        for buffer in self.buffers:
            buffer.current_size = self.ring.current_size
        return [states, rewards, actions, dones, probs]

    def concat_buffer_samples(self):
        return list(self.ring.gather(self.ring.sample_slots()))

    # ------------------------------------------------------------------ returns (acer/agent.py:172-208)
    def calculate_returns(self, rewards, dones, values=None, selected_critic_logits=None, selected_importance=None):
        """Retrace targets [T, E]; `values` has T+1 rows, `dones[t]` is the flag after step t."""
        T, E = rewards.shape
        gate = torch.cat([dones.new_zeros(1, E), dones])          # the kernel's convention: row t+1 gates step t
        return ops.retrace_returns(rewards, gate, values[:T].contiguous(), values[T].contiguous(),
                                   selected_critic_logits.contiguous(), selected_importance.contiguous(), self.gamma)

    # ------------------------------------------------------------------ losses and their output gradients (:210-291)
    def calculate_losses(self, action_probs, values, returns, selected_probs, selected_importance, selected_critic_logits):
        """`values` already without the bootstrap row; all tensors flat over the T*E samples."""
        entropy = (-(action_probs * torch.log(action_probs + self.epsilon)).sum(1)).mean()
        advantages = returns - values
        log_probs = torch.log(selected_probs + self.epsilon)
        action_gain = log_probs * (advantages * torch.clamp(selected_importance, max=self.importance_c)).detach()
        action_loss = -action_gain.mean()
        value_loss = ((returns.detach() - selected_critic_logits) ** 2 * 0.5).mean() * self.value_loss_coef
        if self.trust_region:
            return -(action_loss - self.entropy_coef * entropy) * self.n_steps * self.n_envs, value_loss
        return action_loss + self.value_loss_coef * value_loss - self.entropy_coef * entropy

    def calculate_grads(self, losses, action_probs, critic_logits, avg_action_probs):
        """d(objective)/d(action_probs), d(objective)/d(critic_logits) for the T*E samples -- what the reference's
        tape pushes into the network (`calculate_grads`): plain derivatives without a trust region; with one, the
        policy gradient g is projected so that its component along k = -avg_probs / probs exceeds delta by nothing."""
        if not self.trust_region:
            d_probs, d_critic = torch.autograd.grad(losses, [action_probs, critic_logits], allow_unused=True)
            return d_probs, d_critic if d_critic is not None else torch.zeros_like(critic_logits)
        loss, value_loss = losses
        g, = torch.autograd.grad(loss, [action_probs], retain_graph=True)
        k = -avg_action_probs / (action_probs.detach() + self.epsilon)
        adj = torch.clamp(((k * g).sum(-1) - self.delta) / ((k ** 2).sum(-1) + self.epsilon), min=0.0)
        g = g - adj.view(-1, 1) * k
        d_critic, = torch.autograd.grad(value_loss, [critic_logits])
        return -g / (self.n_envs * self.n_steps), d_critic

    def _avg_outputs(self, states):
        params = {name: self.avg_flat[off:off + int(np.prod(shape))].view(shape) for name, off, shape in self._views}
        with torch.no_grad():
            actor, _ = torch.func.functional_call(self.net.module, params, (self.net.scaled(states),))
        return actor

    def update_avg_weights(self):
        """ema.apply(trainable_variables) + set_weights (acer/agent.py:116-126, 338-339)."""
        if not self._ema_started:
            self.avg_flat.copy_(self.net.flat_param)
            self._ema_started = True
        else:
            self.avg_flat.lerp_(self.net.flat_param, 1.0 - self.ema_alpha)

    def update_gradients(self, states, rewards, actions, dones, previous_action_probs):
        T, E, A = self.n_steps, self.n_envs, self.n_actions
        n = T * E
        flat_states = states.reshape(((T + 1) * E,) + self.input_shape)
        actor_full, critic_full = self.net.forward(flat_states, training=True)      # [(T+1)E, A] each
        critic_full = critic_full.view(-1, A)
        avg_action_probs = self._avg_outputs(flat_states)[:n]
        values = (actor_full * critic_full).sum(-1).view(T + 1, E)
        action_probs = actor_full[:n].clone().requires_grad_(True)                  # clip_last_step: time-major => first T rows
        critic_logits = critic_full[:n].clone().requires_grad_(True)
        index = actions.reshape(n, 1).long()
        selected_probs = action_probs.gather(1, index).squeeze(1)
        selected_critic_logits = critic_logits.gather(1, index).squeeze(1)
        importance_weights = action_probs / (previous_action_probs.reshape(n, A) + self.epsilon)
        selected_importance = importance_weights.gather(1, index).squeeze(1)
        returns = self.calculate_returns(rewards, dones, values, selected_critic_logits.detach().view(T, E),
                                         selected_importance.detach().view(T, E)).reshape(n)
        losses = self.calculate_losses(action_probs, values[:T].reshape(n), returns, selected_probs, selected_importance,
                                       selected_critic_logits)
        d_probs, d_critic = self.calculate_grads(losses, action_probs, critic_logits, avg_action_probs)
        d_actor_full = torch.zeros_like(actor_full)
        d_critic_full = torch.zeros_like(critic_full)
        d_actor_full[:n] = d_probs
        d_critic_full[:n] = d_critic
        self.last_losses = losses
        self.net.backward_and_step(d_actor_full, d_critic_full, self.grad_norm)
        self.update_avg_weights()

    def train_step(self):
        batch = self.get_batch()
        self.buffer_current_size += 1
        self.update_gradients(*batch)
        if self.replay_ratio > 0 and self.buffer_current_size >= self.buffers[0].initial_size:
            for _ in range(self._poisson.poisson(self.replay_ratio)):
                self.update_gradients(*self.concat_buffer_samples())
