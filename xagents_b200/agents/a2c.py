"""A2C with the reference's surface (xagents/a2c/agent.py:9-218), hot path on the device.

The rollout is written step by step into preallocated time-major device buffers (uint8 frames stay
uint8) instead of growing eight Python lists; returns come from `xa_nstep_returns_f32`; the loss and
its gradients w.r.t. the model outputs from `xa_a2c_loss_f32`.  The A2C loss is a mean over all T*E
samples, so the update consumes the buffers in time-major order: no flatten, no gather.
"""
import math

import numpy as np
import torch

from .. import ops
from ..hotpath import A2CHotPath
from .base import OnPolicy
from .models import adapt


class A2C(OnPolicy):
    def __init__(self, envs, model, entropy_coef=0.01, value_loss_coef=0.5, grad_norm=0.5, **kwargs):
        super().__init__(envs, model, **kwargs)
        self.entropy_coef = entropy_coef
        self.value_loss_coef = value_loss_coef
        self.grad_norm = grad_norm
        self.net = adapt(model, self.img_inputs)
        self.output_is_softmax = bool(getattr(self.net, 'output_is_softmax', False))
        self.actor_kind = 'normal' if not self.discrete else ('probs' if self.output_is_softmax else 'logits')
        self.comm = getattr(self.net, 'comm', None)
        self.action_source = None        # tests: callable(step, actor_out) -> actions, to replay a fixed rollout
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(int(self.seed) if self.seed else 0)
        self._rng_offset = 0             # Philox counter offset of the in-kernel sampler
        self.rollout_source = None       # rollouts produced elsewhere (feeds.HostRolloutFeed): callable(agent) that leaves a
        #                                  complete rollout in the ro_* buffers and returns the bootstrap values [E]
        self._fed_last_values = None
        self._a2c_pipeline = None
        # The rollout as ONE CUDA graph (`graph_rollout`): with a device-resident batched environment the n_steps x (policy
        # evaluation, sampler, environment step, row writes) are ~20 launches per step of a few microseconds each -- launch-
        # bound.  'auto' captures them once (Philox offset in device memory, the environment's generator registered with the
        # graph) when everything in the loop lives on the device; False keeps the eager loop.
        self.graph_rollout = 'auto'
        self._rollout_graph = None
        self._alloc_rollout()

    # ------------------------------------------------------------------ buffers
    def _alloc_rollout(self):
        T, E, dev, f32 = self.n_steps, self.n_envs, self.device, torch.float32
        first_dtype = self.states.dtype if self.batched else np.asarray(self.states[0]).dtype
        self.obs_dtype = torch.uint8 if (self.img_inputs and first_dtype in (np.uint8, torch.uint8)) else f32
        self.ro_states = torch.empty((T, E) + self.input_shape, dtype=self.obs_dtype, device=dev)
        self.ro_rewards = torch.empty((T, E), dtype=f32, device=dev)
        self.ro_values = torch.empty((T, E), dtype=f32, device=dev)
        self.ro_dones = torch.empty((T + 1, E), dtype=f32, device=dev)
        self.ro_log_probs = torch.empty((T, E), dtype=f32, device=dev)
        self.ro_entropies = torch.empty((T, E), dtype=f32, device=dev)
        a_shape = (T, E) if self.discrete else (T, E, self.n_actions)
        self.ro_actions = torch.empty(a_shape, dtype=f32, device=dev)
        self.ro_actor = torch.empty((T, E, self.n_actions), dtype=f32, device=dev)
        self.ro_returns = torch.empty((T, E), dtype=f32, device=dev)
        self.loss_scalars = torch.zeros(4, dtype=f32, device=dev)

    def _to_device(self, array, dtype=None):
        if isinstance(array, torch.Tensor):                        # batched device environments hand tensors over
            return array.to(self.device, dtype=dtype)
        t = torch.as_tensor(np.ascontiguousarray(array))
        return t.to(self.device, dtype=dtype, non_blocking=True)

    def _env_actions(self, actions):
        """What `step_envs` gets: the device tensor itself for a batched environment, host values for a list of envs."""
        if self.batched:
            return actions
        env_actions = actions.cpu().numpy()
        return env_actions.astype(np.int64) if self.discrete else env_actions

    # ------------------------------------------------------------------ distribution (a2c/agent.py:50-94)
    def _log_prob_entropy(self, actor_out, actions):
        if not self.discrete:
            k = actor_out.shape[-1]
            log2pi = math.log(2.0 * math.pi)
            logp = -0.5 * ((actions - actor_out) ** 2).sum(-1) - 0.5 * k * log2pi
            return logp, torch.full_like(logp, 0.5 * k * (1.0 + log2pi))
        lsm = torch.log_softmax(torch.log(actor_out) if self.output_is_softmax else actor_out, dim=-1)
        logp = lsm.gather(-1, actions.long().view(-1, 1)).squeeze(-1)
        p = lsm.exp()
        return logp, -torch.where(p > 0, p * lsm, torch.zeros_like(p)).sum(-1)     # multiply_no_nan (tfp Categorical.entropy)

    def sample_actions(self, actor_out):
        if not self.discrete:
            return actor_out + torch.randn(actor_out.shape, device=actor_out.device, generator=self._gen)
        lsm = torch.log_softmax(torch.log(actor_out) if self.output_is_softmax else actor_out, dim=-1)
        u = torch.rand(lsm.shape, device=lsm.device, generator=self._gen).clamp_(1e-12, 1.0)
        return torch.argmax(lsm - torch.log(-torch.log(u)), dim=-1).float()       # Gumbel-max

    def get_model_outputs(self, inputs, models=None, training=True, actions=None, step=None):
        """[actions, log probs, critic output, entropy, actor output] like a2c/agent.py:65-94 (device tensors)."""
        x = inputs if isinstance(inputs, torch.Tensor) else self._to_device(inputs, self.obs_dtype)
        actor_out, critic = self.net.forward(x, training=training)
        if actions is None and self.action_source is None:
            # sample + log_prob + entropy in one kernel (Philox stream keyed by the agent seed)
            rows = None
            if step is not None and getattr(self, '_rollout_rows', False):     # the rollout loop: write row `step` of the buffers directly
                rows = (self.ro_actions[step], self.ro_log_probs[step], self.ro_entropies[step])
            actions, logp, entropy = self._sample(actor_out, rows)
            return actions, logp, critic, entropy, actor_out
        if actions is None:
            actions = self.action_source(step, actor_out)
            actions = actions if isinstance(actions, torch.Tensor) else self._to_device(actions, torch.float32)
        logp, entropy = self._log_prob_entropy(actor_out, actions)
        return actions, logp, critic, entropy, actor_out

    def _sample(self, actor_out, rows=None):
        """sample + log_prob + entropy in one kernel (Philox stream keyed by the agent seed).  Inside a captured rollout the offset
        is `_philox` (device memory) + the steps taken since the capture began; the counter is advanced once, at the rollout's end."""
        capturing = getattr(self.net, 'capturing', False) and getattr(self, '_philox', None) is not None
        if capturing:
            out = ops.policy_step(actor_out, actor_kind=self.actor_kind, counter=self._philox, offset=self._rng_offset - self._philox_base,
                                  advance=0, out=rows, seed=int(self.seed) if self.seed else 0)
        else:
            out = ops.policy_step(actor_out, actor_kind=self.actor_kind, out=rows, seed=int(self.seed) if self.seed else 0, offset=self._rng_offset)
        self._rng_offset += 2 * self.n_actions
        return out

    # ------------------------------------------------------------------ rollout (a2c/agent.py:96-139)
    def get_batch(self):
        """Run the environments for n_steps; returns the time-major device buffers
        [states, rewards, actions, critic_output, dones, log_probs, entropies, actor_output]."""
        if self.rollout_source is not None:
            # the rollout was produced elsewhere (host-side actors, a replay of recorded data): the feed uploads it into
            # the time-major buffers and supplies the bootstrap values that `get_states()` would have been evaluated for
            self._fed_last_values = self.rollout_source(self)
            self.steps += self.n_steps * self.n_envs               # what step_envs counts (base.py:425)
            return [self.ro_states, self.ro_rewards, self.ro_actions, self.ro_values, self.ro_dones, self.ro_log_probs,
                    self.ro_entropies, self.ro_actor]
        if self._rollout_graph is None and self._can_graph_rollout():
            self._capture_rollout()
        if self._rollout_graph:
            self._flush_episode_log()                              # the replay overwrites the rows the pending log entries view
            self._rollout_graph.replay()                           # the whole rollout: one launch
            self.steps += self.n_steps * self.n_envs
            self._episode_log = [(self.ro_dones[t + 1], self._ro_sums[t]) for t in range(self.n_steps)]
        else:
            self._rollout_loop()
        return [self.ro_states, self.ro_rewards, self.ro_actions, self.ro_values, self.ro_dones, self.ro_log_probs,
                self.ro_entropies, self.ro_actor]

    def _fused_rollout_applies(self):
        """Every rollout step as  network -> sampler -> environment  with all three writing the rollout rows in place: needs the
        native batched environment, a model that can write its outputs into given tensors, and the stock step / sampling code."""
        from .base import BaseAgent
        cls = type(self)
        return (self.batched and getattr(self.envs, 'native_step', False) and self.action_source is None and self.obs_dtype == torch.uint8
                and hasattr(self.net, 'infer_into') and cls.get_model_outputs is A2C.get_model_outputs
                and cls.step_envs is BaseAgent.step_envs and cls._rollout_loop is A2C._rollout_loop)

    def _rollout_loop_fused(self, capturing=False):
        """The same rollout (a2c/agent.py:113-139 around base.py:408-426) in three native calls per step: the network writes
        actor / value rows t, the sampler the action / log-prob / entropy rows t, the environment kernel the frame row t + 1
        (the frames a step returns are the next step's input), the reward row t and the done row t + 1, and keeps the episode
        sums -- no copies, no elementwise launches.  Bit-identical buffers to `_rollout_loop` on the same environment."""
        T, envs = self.n_steps, self.envs
        if getattr(self, '_ro_sums', None) is None:
            self._ro_sums = torch.empty((T, self.n_envs), dtype=torch.float32, device=self.device)
        if getattr(self, '_ro_spare', None) is None:
            self._ro_spare = torch.empty_like(self.ro_states[0])
        if not capturing:
            self._flush_episode_log()                              # pending entries view rows this rollout overwrites
        self.ro_states[0].copy_(envs.states)
        self.ro_dones[0].copy_(self.dones)
        for t in range(T):
            self.net.infer_into(self.ro_states[t], self.ro_actor[t], self.ro_values[t])
            self._sample(self.ro_actor[t], (self.ro_actions[t], self.ro_log_probs[t], self.ro_entropies[t]))
            envs.step_into(self.ro_states[t + 1] if t + 1 < T else self._ro_spare, self.ro_rewards[t], self.ro_dones[t + 1],
                           self._episode_sums, self._ro_sums[t])
        self.dones.copy_(self.ro_dones[T])
        self.states = envs.states
        self.steps += T * self.n_envs
        self._episode_log = [(self.ro_dones[t + 1], self._ro_sums[t]) for t in range(T)]

    def _rollout_loop(self, capturing=False):
        """n_steps x (model outputs -> row t of the rollout buffers -> environment step), a2c/agent.py:113-139."""
        if self._fused_rollout_applies():
            return self._rollout_loop_fused(capturing)
        step_states, step_dones = self.get_states(), self.get_dones()
        self._rollout_rows = True                                  # the sampler writes its three rows in place
        for t in range(self.n_steps):
            states_d = self._to_device(step_states, self.obs_dtype)
            actions, logp, values, entropy, actor_out = self.get_model_outputs(states_d, training=False, step=t)
            self.ro_states[t].copy_(states_d)
            for row, value in ((self.ro_actions[t], actions), (self.ro_log_probs[t], logp), (self.ro_entropies[t], entropy)):
                if value.data_ptr() != row.data_ptr():             # an `action_source` / a subclass produced them elsewhere
                    row.copy_(value)
            self.ro_values[t].copy_(values)
            self.ro_dones[t].copy_(self._to_device(step_dones))
            self.ro_actor[t].copy_(actor_out)
            *_, step_rewards, step_dones, step_states = self.step_envs(self._env_actions(actions), True, False)
            self.ro_rewards[t].copy_(self._to_device(step_rewards))
            if capturing:                                          # episode sums at the moment of each done flag
                self._ro_sums[t].copy_(self._episode_log.pop()[1])
        self.ro_dones[self.n_steps].copy_(self._to_device(step_dones))
        self._rollout_rows = False

    # ------------------------------------------------------------------ the rollout as one CUDA graph
    def _can_graph_rollout(self):
        if self.graph_rollout is False or self._rollout_graph is False or self.device.type != 'cuda':
            return False
        from .models import TorchModel
        ok = (self.batched and hasattr(self.envs, 'GRAPH_STATE') and hasattr(self.envs, 'register_graph') and self.action_source is None
              and isinstance(self.net, TorchModel) and (self.graph_rollout is True or type(self).__name__ in ('A2C', 'PPO')))
        if not ok:
            assert self.graph_rollout is not True, 'graph_rollout=True needs a batched device environment and a TorchModel'
            self._rollout_graph = False
        return ok

    def _capture_rollout(self):
        envs, dev = self.envs, self.device
        self._flush_episode_log()
        self._ro_sums = torch.empty((self.n_steps, self.n_envs), dtype=torch.float32, device=dev)
        self._philox = torch.full((1,), self._rng_offset, dtype=torch.int64, device=dev)
        self._philox_base = self._rng_offset
        envs.sync_rng()                                            # the captured steps count from the device counter alone
        # fixed-address homes for everything the loop reads on entry and replaces on exit
        home = {name: getattr(envs, name).clone() for name in envs.GRAPH_STATE}
        home['agent.dones'], home['agent.sums'] = self.dones.clone(), self._episode_sums.clone()

        def enter():                                               # point the live objects at the homes
            for name in envs.GRAPH_STATE:
                setattr(envs, name, home[name])
            self.states, self.dones, self._episode_sums = envs.states, home['agent.dones'], home['agent.sums']

        def leave():                                               # what the loop left behind goes back into the homes
            final = {name: getattr(envs, name) for name in envs.GRAPH_STATE}
            final['agent.dones'], final['agent.sums'] = self.dones, self._episode_sums
            for name, t in final.items():
                if t.data_ptr() != home[name].data_ptr():
                    home[name].copy_(t)
            enter()

        saved = {k: v.clone() for k, v in home.items()}
        gen_state, steps, log, rng_offset = envs.rng_state(), self.steps, list(self._episode_log), self._rng_offset
        self.net.capturing = True                                  # no nested graph replays inside the capture
        stream = torch.cuda.Stream(dev)
        stream.wait_stream(torch.cuda.current_stream(dev))
        try:
            with torch.cuda.stream(stream):
                enter()
                self._rollout_loop(capturing=True)                 # warm-up outside the capture (module loading, allocator, cuDNN)
                stream.synchronize()
                for k, v in saved.items():                         # ... undone: same states, same random streams as before
                    home[k].copy_(v)
                envs.set_rng_state(gen_state)
                self._rng_offset = rng_offset
                self._philox.fill_(rng_offset)
                enter()
                graph = torch.cuda.CUDAGraph()
                envs.register_graph(graph)
                with torch.cuda.graph(graph, stream=stream):
                    self._rollout_loop(capturing=True)
                    ops.bump_u64(self._philox, self._rng_offset - rng_offset)   # the sampler's and the environment's Philox
                    envs.sync_rng()                                             # counters advance once per replay
                    leave()
        finally:
            self.net.capturing = False
        torch.cuda.current_stream(dev).wait_stream(stream)
        self.steps, self._episode_log = steps, log
        self._rollout_graph = graph

    def _bootstrap_values(self):
        if self.rollout_source is not None:
            return self._fed_last_values
        return self.get_model_outputs(self.get_states(), training=False, actions=self.ro_actions[0])[2]

    def calculate_returns(self, rewards, dones, values=None, selected_critic_logits=None, selected_importance=None):
        """n-step returns [T, E] (a2c/agent.py:141-171): bootstrap from the current states, then the scan kernel."""
        rewards = rewards if isinstance(rewards, torch.Tensor) else self._to_device(rewards, torch.float32)
        dones = dones if isinstance(dones, torch.Tensor) else self._to_device(dones, torch.float32)
        out = self.ro_returns if tuple(rewards.shape) == tuple(self.ro_returns.shape) else None
        return ops.nstep_returns(rewards, dones, self._bootstrap_values(), self.gamma, out=out)

    def np_train_step(self):
        states, rewards, actions, critic_output, dones, *_ = self.get_batch()
        returns = self.calculate_returns(rewards, dones)
        return self.concat_step_batches(states, returns, actions, critic_output)

    # ------------------------------------------------------------------ update (a2c/agent.py:190-218)
    def hot_path(self):
        """The agent's A2CHotPath (returns scan + fused loss, launch arguments resolved once) over the `ro_*` buffers."""
        bound = dict(rewards=self.ro_rewards, values=self.ro_values, dones=self.ro_dones, actions=self.ro_actions,
                     returns=self.ro_returns)
        hp = self._a2c_pipeline
        if hp is None:
            hp = self._a2c_pipeline = A2CHotPath(self.n_steps, self.n_envs, self.n_actions, gamma=self.gamma,
                                                 entropy_coef=self.entropy_coef, value_loss_coef=self.value_loss_coef,
                                                 actor_kind=self.actor_kind, device=self.device, buffers=bound)
        else:
            moved = {name: t for name, t in bound.items() if getattr(hp, name).data_ptr() != t.data_ptr()}
            if moved:
                hp.bind(**moved)
        stream = torch.cuda.current_stream(self.device)
        if hp._args is None or hp.stream != stream:
            hp.prepare(stream)
        return hp

    def train_step(self):
        states, rewards, actions, old_values, dones, *_ = self.get_batch()
        returns = self.calculate_returns(rewards, dones)
        if not isinstance(returns, torch.Tensor):
            returns = self._to_device(returns, torch.float32)
        if returns.data_ptr() != self.ro_returns.data_ptr():       # an overriding calculate_returns made its own tensor
            self.ro_returns = returns.contiguous()
        hp = self.hot_path()
        n = self.n_steps * self.n_envs
        forward_into = getattr(self.net, 'forward_into', None)
        flat_states = states.reshape((n,) + self.input_shape)
        if forward_into is not None:
            forward_into(flat_states, hp.actor_out, hp.critic_out, 0)
        else:
            actor_out, critic = self.net.forward(flat_states, training=True)
            hp.actor_out.copy_(actor_out.reshape(n, -1))
            hp.critic_out.copy_(critic.reshape(n))
        hp.run(returns=False)                                      # fused loss forward + backward (a2c/agent.py:202-215)
        self.loss_scalars = hp.scalars
        self.net.backward_and_step(hp.d_actor, hp.d_values, self.grad_norm)
