"""Model adapters: what the agents need from "the model" on either side of the hot path.

The reference hands a compiled `tf.keras.Model` to its agents and lets a GradientTape differentiate the
loss (xagents/ppo/agent.py:112-137).  Here the loss kernel already returns d loss / d(actor_out, critic_out),
so a model only has to (1) run forward on a device batch and (2) back-propagate those two output
gradients and apply its optimiser.  Any object with this interface works:

    forward(states, training=True) -> (actor_out [n, A] fp32, critic_out [n] fp32)   device, DLPack-able
    backward_and_step(d_actor, d_values, grad_norm) -> None
    output_is_softmax : bool      (Categorical(probs=) instead of (logits=), a2c/agent.py:42-43)

`TorchModel` wraps a torch.nn.Module returning (actor_out, critic_out) and keeps every trainable tensor
in ONE flat fp32 buffer so that global-norm clip + Adam is the fused two-launch step of csrc/optim.cu
(Keras Adam semantics, utils/common.py:476; defaults utils/cli.py:14-25) and the gradient all-reduce is
a single NCCL call.  `KerasModel` drives a tf.keras.Model through DLPack (untestable in this image:
TensorFlow is not installable here; it is imported lazily).
"""
import torch

from .. import ops
from ..optim import FlatAdam


class TorchModel:
    def __init__(self, module, lr=7e-4, beta1=0.9, beta2=0.999, epsilon=1e-7, img_inputs=None, output_is_softmax=False,
                 comm=None, tensor_core_inference=False, graph_inference=True, role='actor_critic', fused_c1='auto', native_plan=True):
        """`fused_c1` (sharded training, comm.world_size > 1): 'auto' / True = the gradient all-reduce, the global-norm clip
        and Adam run as ONE kernel over NVLink peer memory (xagents_b200/peer.py) when the box can map peer memory -- the flat
        parameter and gradient buffers then live in peer-mapped memory and Adam's m, v exist only for this rank's shard;
        False = NCCL all-reduce + the two-launch clip+Adam.
        `native_plan`: a module that offers `plan()` (NatureCnnTc) trains through its two native calls per minibatch, the backward
        writing the flat gradient buffer directly (agents/tc_plan.py); False keeps torch autograd around the same kernels."""
        assert role in ('actor_critic', 'actor', 'critic'), f'unknown model role `{role}`'
        self.module = module
        self.role = role                                          # 'actor' / 'critic': TRPO's separate single-output networks
        self.output_is_softmax = output_is_softmax
        self.img_inputs = img_inputs
        self.lr, self.beta1, self.beta2, self.epsilon = lr, beta1, beta2, epsilon
        self.comm = comm
        self.step = 0
        params = [p for p in module.parameters() if p.requires_grad]
        assert params, 'model has no trainable parameters'
        dev = params[0].device
        assert dev.type == 'cuda', 'the model must live on the GPU: xagents_b200 has no CPU path'
        n = sum(p.numel() for p in params)
        pad = (-n) % 4                                            # float4 path of the optimiser kernel
        self.fused = None
        if fused_c1 and comm is not None and comm.world_size > 1:
            from .. import peer
            if peer.available(comm):
                self.fused = peer.FusedAllReduceAdam(comm, n, lr=lr, beta1=beta1, beta2=beta2, eps=epsilon)
            else:
                assert fused_c1 == 'auto', f'no peer-memory transport on this box: {peer.probe_errors()}'
        if self.fused is not None:
            self.flat_param, self.flat_grad = self.fused.param, self.fused.grad      # peer-mapped, padded to world * shard
            self.optimizer, self.m, self.v, self.workspace = None, self.fused.m, self.fused.v, None
        else:
            self.flat_param = torch.zeros(n + pad, dtype=torch.float32, device=dev)
            self.flat_grad = torch.zeros_like(self.flat_param)
            self.optimizer = FlatAdam(self.flat_param, self.flat_grad, lr=lr, beta1=beta1, beta2=beta2, eps=epsilon)
            self.m, self.v, self.workspace = self.optimizer.m, self.optimizer.v, self.optimizer.workspace
        off = 0
        for p in params:                                          # re-seat parameters and grads as views
            k = p.numel()
            self.flat_param[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + k].view_as(p)
            p.grad = self.flat_grad[off:off + k].view_as(p)
            off += k
        self.n_params = n
        self._outputs = None
        self._native_plan = bool(native_plan) and role == 'actor_critic' and hasattr(module, 'plan')
        self._refreshable = [m for m in module.modules() if hasattr(m, 'refresh')]
        for mod in self._refreshable:
            mod.refresh()
        # rollout-time policy evaluation (no grad) through the tcgen05 convolution + GEMM pipeline
        self._tc_forward = None
        self.capturing = False                                    # set by a caller that is capturing a CUDA graph around forward()
        self._graph_inference = graph_inference
        if tensor_core_inference or getattr(module, 'takes_uint8', False):
            from .tc_conv import NatureCnnTcForward
            assert isinstance(module, NatureCNN), 'tensor-core inference is implemented for NatureCNN'
            self._tc_forward = NatureCnnTcForward(module)
            self._refreshable.append(self._tc_forward)

    def forward(self, states, training=True):
        if not training and self._tc_forward is not None and states.dtype == torch.uint8:
            if self._graph_inference and not self.capturing:      # fixed rollout batch: one graph replay per step
                a, c = self._tc_forward.graphed(states.shape[0])(states)
                return a.clone(), c.clone()
            a, c = self._tc_forward(states)                       # the plan's own buffers: handed out as they are only inside a
            return (a, c) if self.capturing else (a.clone(), c.clone())   # captured rollout, whose next kernels consume them
        x = self.scaled(states)
        if training and self._native_plan and x.dtype in (torch.uint8, torch.bfloat16):
            plan = self.module.plan(x.shape[0], self.flat_param, self.flat_grad.numel())
            self._outputs = plan
            return plan.forward(x.contiguous())
        with torch.set_grad_enabled(training):
            out = self.module(x)
        if self.role == 'actor_critic':
            actor, critic = out
        else:
            out = out[0] if isinstance(out, (tuple, list)) else out
            actor, critic = (out, None) if self.role == 'actor' else (None, out)
        if critic is not None:
            critic = critic.reshape(-1)                           # tf.squeeze, a2c/agent.py:84
        self._outputs = (actor, critic) if training else None
        return (None if actor is None else actor.detach().contiguous(),
                None if critic is None else critic.detach().contiguous())

    def infer_into(self, states, actor_dst, critic_dst):
        """Rollout-time forward (training=False) with the outputs written into the caller's tensors -- rows of the rollout
        buffers: in place through the tensor-core inference plan, forward() + two copies otherwise."""
        if self._tc_forward is not None and states.dtype == torch.uint8:
            self._tc_forward(states, out=(actor_dst, critic_dst))
            return
        actor, critic = self.forward(states, training=False)
        actor_dst.copy_(actor.reshape(actor_dst.shape))
        critic_dst.copy_(critic.reshape(critic_dst.shape))

    @property
    def reads_through_permutation(self):
        """True when `forward_into(store, ..., idx=ids)` evaluates the minibatch store[ids] without a gathered copy of it (the
        native plan's first layer fetches every frame by its id): the prepared PPO pipeline then skips its frame gather."""
        return self._native_plan and getattr(self.module, 'takes_uint8', False)

    def forward_into(self, states, actor_dst, critic_dst, i=0, idx=None, time_major=None):
        """Training forward with the outputs written into the caller's tensors (the prepared pipelines' per-minibatch output
        rows): no copies with the native plan, forward() + two copies otherwise.  `idx` (int32 [n], device; needs
        `reads_through_permutation`): the minibatch is states[idx] -- env-major sample ids of the time-major rollout
        `states` [T*E, ...] when `time_major=(T, E)` -- fetched by the first layer itself."""
        if idx is not None:
            assert self.reads_through_permutation and states.dtype == torch.uint8 and states.is_contiguous()
            plan = self.module.plan(idx.numel(), self.flat_param, self.flat_grad.numel())
            self._outputs = plan
            plan.forward(states, out=(actor_dst, critic_dst), idx=idx, time_major=time_major)
            return
        x = self.scaled(states)
        if self._native_plan and x.dtype in (torch.uint8, torch.bfloat16):
            plan = self.module.plan(x.shape[0], self.flat_param, self.flat_grad.numel())
            self._outputs = plan
            plan.forward(x if x.is_contiguous() else x.contiguous(), out=(actor_dst, critic_dst))
            return
        actor, critic = self.forward(states, training=True)
        if actor is not None:
            actor_dst.copy_(actor.reshape(actor_dst.shape))
        if critic is not None:
            critic_dst.copy_(critic.reshape(critic_dst.shape))

    def scaled(self, states):
        """The network's input for a batch of stored observations: fp32, images divided by 255 (base.py:505-506)."""
        x = states
        if getattr(self.module, 'takes_uint8', False) and x.dtype in (torch.uint8, torch.bfloat16):
            return x                                              # the tensor-core network scales inside its first kernel
        scale = self.img_inputs if self.img_inputs is not None else (x.dtype == torch.uint8)
        if x.dtype != torch.float32:
            x = x.float()
        return x / 255.0 if scale else x

    def backward_and_step(self, d_actor, d_values, grad_norm=None):
        if self._native_plan and not isinstance(self._outputs, tuple):
            self._outputs.backward(d_actor, d_values, self.flat_grad)      # writes every element: no zero-fill
        else:
            actor, critic = self._outputs
            self.flat_grad.zero_()
            pairs = [(o, g) for o, g in ((actor, d_actor), (critic, d_values)) if o is not None]
            torch.autograd.backward([o for o, _ in pairs], [g.view_as(o) for o, g in pairs])
        self._outputs = None
        self.step += 1
        if self.fused is not None:                                # collective C1 + clip + Adam: one kernel over peer memory
            self.fused.lr = self.lr
            self.fused.step(self.step, grad_norm)
            for mod in self._refreshable:
                mod.refresh()
            return
        scale = 1.0
        if self.comm is not None and self.comm.world_size > 1:
            self.comm.all_reduce_gradients_async(self.flat_grad)  # collective C1 (NCCL, comm stream) ...
            self.comm.wait_gradients()                            # ... which the optimiser has to wait for
            scale = 1.0 / self.comm.world_size
        opt = self.optimizer
        opt.t, opt.lr = self.step - 1, self.lr                    # `step` and `lr` stay the attributes callers may set
        opt.step(grad_norm, scale)
        for mod in self._refreshable:                             # bf16 operand copies of tensor-core layers
            mod.refresh()


class KerasModel:
    """tf.keras.Model adapter (DLPack both ways).  Needs TensorFlow with GPU support at call time."""

    def __init__(self, model, img_inputs=False):
        import tensorflow as tf                                   # lazily: absent from this image
        self.tf, self.model, self.img_inputs = tf, model, img_inputs
        acts = [getattr(layer, 'activation', None) for layer in model.layers[-2:]]
        self.output_is_softmax = tf.keras.activations.softmax in acts
        self._tape = self._outputs = None

    def forward(self, states, training=True):
        tf = self.tf
        x = tf.experimental.dlpack.from_dlpack(states.__dlpack__())
        if self.img_inputs:
            x = tf.cast(x, tf.float32) / 255.0
        self._tape = tf.GradientTape() if training else None
        if training:
            with self._tape:
                actor, critic = self.model(x, training=True)
                critic = tf.squeeze(critic)
            self._outputs = (actor, critic)
        else:
            actor, critic = self.model(x, training=False)
            critic = tf.squeeze(critic)
        return (torch.from_dlpack(tf.experimental.dlpack.to_dlpack(actor)),
                torch.from_dlpack(tf.experimental.dlpack.to_dlpack(critic)))

    def backward_and_step(self, d_actor, d_values, grad_norm=None):
        tf = self.tf
        grads = self._tape.gradient(list(self._outputs), self.model.trainable_variables,
                                    output_gradients=[tf.experimental.dlpack.from_dlpack(d_actor.__dlpack__()),
                                                      tf.experimental.dlpack.from_dlpack(d_values.__dlpack__())])
        if grad_norm is not None:
            grads, _ = tf.clip_by_global_norm(grads, grad_norm)
        self.model.optimizer.apply_gradients(zip(grads, self.model.trainable_variables))
        self._tape = self._outputs = None


def adapt(model, img_inputs):
    """Accept an adapter, a torch module, or a Keras model, as the reference accepts `model`."""
    if hasattr(model, 'forward') and hasattr(model, 'backward_and_step'):
        return model
    if isinstance(model, torch.nn.Module):
        return TorchModel(model, img_inputs=img_inputs)
    if hasattr(model, 'trainable_variables') and hasattr(model, 'layers'):
        assert len(model.layers) > 2, f'Expected a model that has at least 3 layers, got {len(model.layers)}'
        return KerasModel(model, img_inputs=img_inputs)
    raise TypeError(f'unsupported model type {type(model).__name__}: need forward()/backward_and_step(), '
                    f'a torch.nn.Module or a tf.keras.Model')


class NatureCNN(torch.nn.Module):
    """The documented PPO/A2C CNN (README.md:243-259; ppo/models/cnn-actor-critic.cfg:1-42 read as Conv2D):
    32x8/4 -> 64x4/2 -> 64x3/1 -> FC512 shared trunk -> actor (A) and critic (1) heads, orthogonal init."""

    def __init__(self, in_channels=4, n_actions=6, tensor_core_dense=False):
        super().__init__()
        nn = torch.nn
        if tensor_core_dense:                                     # FC512 + heads on the tcgen05 GEMM (tc_dense.py)
            from .tc_dense import TcLinear
            fc = [TcLinear(3136, 512, relu=True)]
            self.actor, self.critic = TcLinear(512, n_actions), TcLinear(512, 1)
        else:
            fc = [nn.Linear(3136, 512), nn.ReLU()]
            self.actor, self.critic = nn.Linear(512, n_actions), nn.Linear(512, 1)
        self.trunk = nn.Sequential(nn.Conv2d(in_channels, 32, 8, 4), nn.ReLU(), nn.Conv2d(32, 64, 4, 2), nn.ReLU(),
                                   nn.Conv2d(64, 64, 3, 1), nn.ReLU(), nn.Flatten(), *fc)
        for mod, gain in [(m, 2 ** 0.5) for m in self.trunk if hasattr(m, 'weight')] + [(self.actor, 0.01), (self.critic, 1.0)]:
            nn.init.orthogonal_(mod.weight, gain)
            nn.init.zeros_(mod.bias)

    def forward(self, x):                                          # x: [n, 84, 84, C] channels-last like the reference
        h = self.trunk(x.permute(0, 3, 1, 2))
        return self.actor(h), self.critic(h)


class AsCodedConv1dCNN(torch.nn.Module):
    """The same `.cfg` as the reference's ModelReader ACTUALLY builds it: the conv sections become
    `Conv1D` layers (utils/common.py:17,231-237), which Keras applies to the 4-D frame batch with the image rows as an
    extended batch dimension -- 1-D convolutions along the width, 32x8/4 -> 64x4/2 -> 64x3/1 on [n, 84, 84, C], then
    flatten (84*7*64 = 37 632) -> FC512 -> heads: 19.3 M parameters against the documented Conv2D network's 1.69 M.
    The 1-D convolutions (4/32/64 channels, a few MFLOP per frame) run through the framework's library path; the FC
    layer that holds 99.6 % of the parameters and most of the flops runs on the tcgen05 GEMM (`TcLinear`) when
    `tensor_core_dense` is set.  Both architectures go through the same agents, loss kernels and fused optimiser."""

    def __init__(self, in_channels=4, n_actions=6, tensor_core_dense=False):
        super().__init__()
        nn = torch.nn
        if tensor_core_dense:
            from .tc_dense import TcLinear
            fc = [TcLinear(84 * 7 * 64, 512, relu=True)]
            self.actor, self.critic = TcLinear(512, n_actions), TcLinear(512, 1)
        else:
            fc = [nn.Linear(84 * 7 * 64, 512), nn.ReLU()]
            self.actor, self.critic = nn.Linear(512, n_actions), nn.Linear(512, 1)
        conv = lambda cin, cout, k, s: nn.Conv2d(cin, cout, (1, k), (1, s))       # Conv1D over an extended batch of rows
        self.trunk = nn.Sequential(conv(in_channels, 32, 8, 4), nn.ReLU(), conv(32, 64, 4, 2), nn.ReLU(), conv(64, 64, 3, 1), nn.ReLU())
        self.fc = nn.Sequential(*fc)
        mods = [m for m in self.trunk if hasattr(m, 'weight')] + [m for m in self.fc if hasattr(m, 'weight')]
        for mod, gain in [(m, 2 ** 0.5) for m in mods] + [(self.actor, 0.01), (self.critic, 1.0)]:
            nn.init.orthogonal_(mod.weight, gain)
            nn.init.zeros_(mod.bias)

    def forward(self, x):                                          # x: [n, 84, 84, C] channels-last like the reference
        h = self.trunk(x.permute(0, 3, 1, 2))                      # [n, 64, 84, 7]
        h = h.permute(0, 2, 3, 1).reshape(h.shape[0], -1)          # Keras flatten order (row, position, channel)
        h = self.fc(h)
        return self.actor(h), self.critic(h)
