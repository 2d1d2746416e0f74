"""TRPO with the reference's surface (xagents/trpo/agent.py:6-348) on top of the PPO hot path.

What TRPO shares with PPO runs on the same kernels: the rollout into time-major device buffers, GAE
(`xa_gae_f32`), the env-major views, whole-batch advantage normalisation (`xa_adv_moments_f32` +
`xa_normalize_adv_f32` with eps = 0: trpo/agent.py:316-319 divides by the bare population std), the permute-gather
of the critic's minibatches (`get_mini_batches`, trpo/agent.py:292-293), the strided `states[::fvp_n_steps]`
subsample (a gather with ids 0, k, 2k, ... in env-major order) and the critic's Adam step (`xa_clip_adam_f32`
without a clip).  What is TRPO's own -- the surrogate gradient, the Fisher-vector products (a Hessian-vector
product of the KL divergence: double back-propagation through the actor), conjugate gradients and the backtracking
line search -- is driver arithmetic on the actor's flat parameter vector and goes through autograd, as it goes
through nested GradientTapes in the reference.  The actor and critic are separate single-output networks
(`TorchModel(role='actor' | 'critic')`).
"""
import numpy as np
import torch

from .. import ops
from .models import TorchModel
from .ppo import PPO


class ActorCriticPair:
    """Rollout-time view of TRPO's two networks as the one actor-critic model the A2C rollout loop drives
    (get_model_outputs(states, [actor, critic]), a2c/agent.py:65-94)."""
    comm = None

    def __init__(self, actor, critic):
        self.actor, self.critic = actor, critic
        self.output_is_softmax = bool(getattr(actor, 'output_is_softmax', False))

    def forward(self, states, training=True):
        assert not training, 'TRPO updates its two networks separately'
        return self.actor.forward(states, training=False)[0], self.critic.forward(states, training=False)[1]

    def backward_and_step(self, *args, **kwargs):
        raise NotImplementedError('TRPO updates its two networks separately')


def _with_role(model, role):
    """-> (adapter, created here).  Raw torch modules are wrapped; adapters must already carry the role."""
    if isinstance(model, TorchModel):
        assert model.role == role, f'Expected a TorchModel with role `{role}`, got `{model.role}`'
        return model, False
    assert isinstance(model, torch.nn.Module), f'TRPO needs torch modules or TorchModel adapters, got {type(model).__name__}'
    return TorchModel(model, role=role), True


class TRPO(PPO):
    def __init__(self, envs, actor_model, critic_model, max_kl=1e-3, cg_iterations=10, cg_residual_tolerance=1e-10,
                 cg_damping=1e-3, actor_iterations=10, critic_iterations=3, fvp_n_steps=5, **kwargs):
        (actor, made_actor), (critic, made_critic) = _with_role(actor_model, 'actor'), _with_role(critic_model, 'critic')
        assert not actor._refreshable, ('TRPO differentiates the actor twice (Fisher-vector products); tensor-core layers '
                                        'implement first derivatives only -- build the actor without tensor_core_dense')
        super().__init__(envs, ActorCriticPair(actor, critic), **kwargs)
        self.model = self.actor = actor
        self.critic = critic
        self.output_models = [actor, critic]
        for net, made in ((actor, made_actor), (critic, made_critic)):
            if made:                                               # like adapt(): raw modules take the agent's image flag
                net.img_inputs = self.img_inputs
        # the reference keeps a cloned Keras model and copies the weights over at every step start
        # (trpo/agent.py:50, 225-233); a copy of the flat parameter vector, applied functionally, is the same thing
        self.old_actor_flat = actor.flat_param.clone()
        self._param_views = self._views(actor)
        self.cg_iterations = cg_iterations
        self.cg_residual_tolerance = cg_residual_tolerance
        self.cg_damping = cg_damping
        self.max_kl = max_kl
        self.critic_iterations = critic_iterations
        self.actor_iterations = actor_iterations
        self.fvp_n_steps = fvp_n_steps
        self.last_update = {}            # diagnostics of the latest actor update (device scalars fetched once)

    # ------------------------------------------------------------------ flat <-> weights (trpo/agent.py:61-118)
    @staticmethod
    def _views(net):
        """(name, offset, shape) of every trainable tensor inside the adapter's flat parameter buffer."""
        out, off = [], 0
        for name, p in net.module.named_parameters():
            if p.requires_grad:
                out.append((name, off, tuple(p.shape)))
                off += p.numel()
        return out

    def flat_to_weights(self, flat, trainable_variables=None, in_place=False):
        if in_place:
            self.actor.flat_param[:self.actor.n_params].copy_(flat[:self.actor.n_params])
            return []
        return [flat[off:off + int(np.prod(shape))].view(shape) for _, off, shape in self._param_views]

    @staticmethod
    def weights_to_flat(to_flatten, trainable_variables=None):
        if not trainable_variables:
            return torch.cat([t.reshape(-1) for t in to_flatten])
        return torch.cat([(t if t is not None else torch.zeros_like(v)).reshape(-1)
                          for t, v in zip(to_flatten, trainable_variables)])

    def _actor_params(self):
        return [p for p in self.actor.module.parameters() if p.requires_grad]

    def _old_actor_output(self, x):
        params = {name: self.old_actor_flat[off:off + int(np.prod(shape))].view(shape) for name, off, shape in self._param_views}
        with torch.no_grad():
            out = torch.func.functional_call(self.actor.module, params, (x,))
        return out[0] if isinstance(out, (tuple, list)) else out

    def _new_actor_output(self, x):
        out = self.actor.module(x)
        return out[0] if isinstance(out, (tuple, list)) else out

    # ------------------------------------------------------------------ KL, losses, FVP, CG (trpo/agent.py:120-223)
    def _log_softmax(self, actor_out):
        return torch.log_softmax(torch.log(actor_out) if self.output_is_softmax else actor_out, dim=-1)

    def calculate_kl_divergence(self, states):
        """mean KL(old || new) over `states`, plus both actor outputs (the reference returns the distributions)."""
        x = self.actor.scaled(states)
        old, new = self._old_actor_output(x), self._new_actor_output(x)
        if self.discrete:
            lo, ln = self._log_softmax(old), self._log_softmax(new)
            kl = (lo.exp() * (lo - ln)).sum(-1)
        else:                                                      # MultivariateNormalDiag, identity scale (a2c/agent.py:59-60)
            kl = 0.5 * ((old - new) ** 2).sum(-1)
        return kl.mean(), old, new

    def calculate_losses(self, states, actions, advantages):
        kl_divergence, old, new = self.calculate_kl_divergence(states)
        new_logp, new_entropy = self._log_prob_entropy(new, actions)
        old_logp, _ = self._log_prob_entropy(old, actions)
        ratio = torch.exp(new_logp - old_logp)
        surrogate_loss = (ratio * advantages).mean() + self.entropy_coef * new_entropy.mean()
        return surrogate_loss, kl_divergence

    def calculate_fvp(self, flat_tangent, states):
        params = self._actor_params()
        kl_divergence, *_ = self.calculate_kl_divergence(states)
        kl_grads = torch.autograd.grad(kl_divergence, params, create_graph=True, allow_unused=True)
        tangents = self.flat_to_weights(flat_tangent)
        gvp = sum((g * t).sum() for g, t in zip(kl_grads, tangents) if g is not None)
        hessian_products = torch.autograd.grad(gvp, params, allow_unused=True)
        return self.weights_to_flat(hessian_products, params) + self.cg_damping * flat_tangent

    def conjugate_gradients(self, flat_grads, states):
        p, r, x = flat_grads.clone(), flat_grads.clone(), torch.zeros_like(flat_grads)
        r_dot_r = torch.dot(r, r)
        iterations = 0
        while iterations < self.cg_iterations and float(r_dot_r) > self.cg_residual_tolerance:
            z = self.calculate_fvp(p, states)
            v = r_dot_r / torch.dot(p, z)
            x += v * p
            r -= v * z
            new_r_dot_r = torch.dot(r, r)
            p = r + (new_r_dot_r / r_dot_r) * p
            r_dot_r = new_r_dot_r
            iterations += 1
        return x

    # ------------------------------------------------------------------ updates (trpo/agent.py:225-297)
    def at_step_start(self):
        self.old_actor_flat.copy_(self.actor.flat_param)

    def update_actor_weights(self, flat_weights, full_step, surrogate_loss, states, actions, advantages):
        learning_rate, accepted = 1.0, False
        losses = [float('nan'), float('nan'), float(surrogate_loss.detach())]
        for _ in range(self.actor_iterations):
            self.flat_to_weights(flat_weights + full_step * learning_rate, in_place=True)
            with torch.no_grad():
                new_surrogate_loss, new_kl_divergence = self.calculate_losses(states, actions, advantages)
            losses = torch.stack([new_surrogate_loss, new_kl_divergence, surrogate_loss.detach()]).tolist()
            improvement = losses[0] - losses[2]
            if np.isfinite(losses[:2]).all() and losses[1] <= self.max_kl * 1.5 and improvement > 0:
                accepted = True
                break
            learning_rate *= 0.5
        else:
            self.flat_to_weights(flat_weights, in_place=True)
        self.last_update = {'accepted': accepted, 'step_scale': learning_rate if accepted else 0.0,
                            'surrogate_before': losses[2], 'surrogate_after': losses[0], 'kl_after': losses[1]}

    def update_critic_weights(self, states, returns):
        """critic_iterations x (ppo_epochs x mini_batches shuffled minibatches) of Adam on mean((V - R)^2)."""
        for _ in range(self.critic_iterations):
            for states_mb, returns_mb in self.get_mini_batches(states, returns):
                _, values = self.critic.forward(states_mb, training=True)
                d_values = (values - returns_mb) * (2.0 / values.shape[0])
                self.critic.backward_and_step(None, d_values, None)

    def train_step(self):
        states, actions, returns, values, _ = self.get_batch()
        T, E, n = self.n_steps, self.n_envs, self.batch_size
        ids = torch.arange(n, dtype=torch.int32, device=self.device)
        # (adv - mean) / std over the whole batch, env-major like the reference's flat tensors
        advantages = ops.normalize_advantages(returns.tensor, values.tensor, 0.0, idx=ids, time_major=(T, E))
        flat_states = states.materialize()
        flat_actions = actions.materialize()
        fvp_states = ops.gather_rows(states.tensor, ids[::self.fvp_n_steps].contiguous(), time_major=(T, E))
        params = self._actor_params()
        surrogate_loss, _ = self.calculate_losses(flat_states, flat_actions, advantages)
        flat_grads = self.weights_to_flat(torch.autograd.grad(surrogate_loss, params, allow_unused=True), params)
        step_direction = self.conjugate_gradients(flat_grads, fvp_states)
        shs = 0.5 * torch.dot(step_direction, self.calculate_fvp(step_direction, fvp_states))
        full_step = step_direction / torch.sqrt(shs / self.max_kl)
        pre_actor_weights = self.actor.flat_param[:self.actor.n_params].clone()
        self.update_actor_weights(pre_actor_weights, full_step, surrogate_loss, flat_states, flat_actions, advantages)
        self.update_critic_weights(states, returns)
