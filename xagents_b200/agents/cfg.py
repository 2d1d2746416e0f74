"""`.cfg`-described networks, as the reference's `ModelReader` reads them (xagents/utils/common.py:167-290),
built as torch modules for the device path.

The file format is the reference's (configparser sections, in order):

    [convolutional-N]  filters, size, stride, activation, initializer, gain
    [flatten-N]
    [dense-N]          units (absent -> next entry of `output_units`), activation, initializer, gain,
                       common=1 (every later dense layer reads THIS layer's output), output=1 (model output)

so the `.cfg` files shipped with the reference load unchanged.  Two readings of a convolutional section exist:

* `conv_dims=1` (default, "as coded"): `ModelReader.create_convolution` instantiates `Conv1D`
  (common.py:17,218-237).  Keras applies it to a [B, H, W, C] frame batch with the image rows as an extended
  batch, i.e. a 1-D convolution along W only; cnn-actor-critic.cfg then flattens 84*7*64 = 37 632 features
  (19 293 351 parameters at 4 input channels and 6 actions).
* `conv_dims=2`: the Conv2D network the reference's README documents (README.md:243-259; 1 687 719 parameters).

Activations stay channels-last between layers, so `[flatten]` yields Keras' feature order and weights are
interchangeable with a Keras model of the same `.cfg` after the usual kernel transposes.  Dense layers run on the
tcgen05 GEMM (`TcLinear`) when `tensor_core_dense` is set and the layer qualifies (in_features % 8 == 0,
activation relu or none); everything else goes through the framework's library path.  The hot path either side
of the network (rollout, returns, gathers, losses, clip+Adam) does not depend on which reading is used.
"""
import configparser
import math

import torch

_ACTIVATIONS = {
    None: None, '': None, 'linear': None,
    'relu': torch.nn.ReLU, 'tanh': torch.nn.Tanh, 'sigmoid': torch.nn.Sigmoid,
    'softmax': lambda: torch.nn.Softmax(dim=-1), 'elu': torch.nn.ELU, 'selu': torch.nn.SELU,
}


def _activation(name):
    assert name in _ACTIVATIONS, f'Unsupported activation `{name}`'
    make = _ACTIVATIONS[name]
    return make() if make else None


class _ConvChannelsLast(torch.nn.Module):
    """Keras-convention convolution on channels-last activations, `valid` padding.

    dims=1: x [..., L, C] -> [..., L', filters]; every leading axis is batch (Keras' Conv1D on a rank>3 input).
    dims=2: x [B, H, W, C] -> [B, H', W', filters]."""

    def __init__(self, dims, in_channels, filters, size, stride, activation):
        super().__init__()
        assert dims in (1, 2)
        self.dims = dims
        conv = torch.nn.Conv1d if dims == 1 else torch.nn.Conv2d
        self.conv = conv(in_channels, filters, size, stride)
        self.act = _activation(activation)

    @property
    def weight(self):
        return self.conv.weight

    @property
    def bias(self):
        return self.conv.bias

    def forward(self, x):
        if self.dims == 1:
            lead = x.shape[:-2]
            y = self.conv(x.reshape((-1,) + x.shape[-2:]).transpose(1, 2)).transpose(1, 2)
            y = y.reshape(lead + y.shape[-2:])
        else:
            y = self.conv(x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
        return self.act(y) if self.act else y


class _Flatten(torch.nn.Module):
    def forward(self, x):
        return x.reshape(x.shape[0], -1)


class _Dense(torch.nn.Module):
    def __init__(self, in_features, units, activation, tensor_core):
        super().__init__()
        use_tc = tensor_core and in_features % 8 == 0 and activation in (None, '', 'linear', 'relu')
        if use_tc:
            from .tc_dense import TcLinear
            self.linear = TcLinear(in_features, units, relu=(activation == 'relu'))
            self.act = None
        else:
            self.linear = torch.nn.Linear(in_features, units)
            self.act = _activation(activation)
        self.on_tensor_cores = use_tc

    @property
    def weight(self):
        return self.linear.weight

    @property
    def bias(self):
        return self.linear.bias

    def forward(self, x):
        y = self.linear(x)
        return self.act(y) if self.act else y


class CfgNetwork(torch.nn.Module):
    """The graph `ModelReader.build_model` wires (common.py:257-290): layers in file order; a dense layer reads the
    `common` layer's output once one has been declared, else the running output; `output=1` layers are the model's
    outputs, in file order.  `forward` returns the single output, or a tuple (actor, critic, ...) of them."""

    def __init__(self, layers, plan, common, outputs, input_shape, activations):
        super().__init__()
        self.layers = torch.nn.ModuleList(layers)
        self.plan = plan                  # per layer: True -> input is the latest `common` layer's output
        self.common_indices = frozenset(common)
        self.output_indices = outputs
        self.input_shape = tuple(input_shape)
        self.layer_activations = activations
        # a2c/agent.py:42-43, 61-63: Categorical(probs=...) when one of the last two layers ends in a softmax
        self.output_is_softmax = 'softmax' in activations[-2:]

    def forward(self, x):
        results, current, common = {}, x, None
        for i, (layer, from_common) in enumerate(zip(self.layers, self.plan)):
            current = layer(common if from_common else current)
            if i in self.common_indices:
                common = current
            if i in self.output_indices:
                results[i] = current
        outs = tuple(results[i] for i in self.output_indices)
        return outs[0] if len(outs) == 1 else outs


class ModelReader:
    """Same constructor and `build_model()` as the reference's reader (common.py:172-196, 257-290).  `optimizer`
    is a dict of Adam keywords (`learning_rate`, `beta_1`, `beta_2`, `epsilon`: what `create_agent` collects,
    common.py:589-594) and is returned with the network by `build_adapter()`; `seed` seeds the initialisers
    (a seed implies glorot_uniform where a section names no initialiser, common.py:207-216)."""

    def __init__(self, cfg_file, output_units, input_shape, optimizer=None, seed=None, conv_dims=1,
                 tensor_core_dense=False):
        self.initializers = {'orthogonal': self._orthogonal, 'glorot_uniform': self._glorot_uniform}
        self.cfg_file = cfg_file
        with open(cfg_file) as cfg:
            self.parser = configparser.ConfigParser()
            self.parser.read_file(cfg)
        self.optimizer = optimizer
        self.output_units = list(output_units)
        self.input_shape = tuple(int(d) for d in (input_shape if hasattr(input_shape, '__len__') else (input_shape,)))
        self.seed = seed
        self.output_count = 0
        self.conv_dims = conv_dims
        self.tensor_core_dense = tensor_core_dense
        self._generator = None

    # ------------------------------------------------------------------ initialisers
    def _orthogonal(self, weight, gain=1.0):
        torch.nn.init.orthogonal_(weight, gain, generator=self._generator)

    def _glorot_uniform(self, weight, gain=1.0):
        torch.nn.init.xavier_uniform_(weight, generator=self._generator)

    def get_initializer(self, section):
        """-> callable(weight) or None (framework default, like Keras' when no initialiser is configured)."""
        name = self.parser[section].get('initializer')
        gain = self.parser[section].get('gain')
        if self.seed is not None:
            name = name or 'glorot_uniform'
        init = self.initializers.get(name)
        if init is None:
            return None
        gain = float(gain) if gain else 1.0
        return lambda weight: init(weight, gain)

    def _finish(self, layer, section):
        init = self.get_initializer(section)
        with torch.no_grad():
            if init is not None:
                init(layer.weight)
            else:                                                  # Keras default: glorot_uniform
                torch.nn.init.xavier_uniform_(layer.weight, generator=self._generator)
            layer.bias.zero_()                                     # Keras default bias initialiser
        return layer

    # ------------------------------------------------------------------ layers
    def create_convolution(self, section, shape):
        sec = self.parser[section]
        filters, size, stride = int(sec['filters']), int(sec['size']), int(sec['stride'])
        assert len(shape) >= 2, f'{section}: a convolution needs [..., length, channels] inputs, got {shape}'
        if self.conv_dims == 2:
            assert len(shape) == 3, f'{section}: conv_dims=2 needs [H, W, C] inputs, got {shape}'
        layer = _ConvChannelsLast(self.conv_dims, shape[-1], filters, size, stride, sec.get('activation'))
        out = lambda n: (n - size) // stride + 1
        if self.conv_dims == 1:
            new_shape = shape[:-2] + (out(shape[-2]), filters)
        else:
            new_shape = (out(shape[0]), out(shape[1]), filters)
        assert min(new_shape) > 0, f'{section}: input {shape} is smaller than the kernel'
        return self._finish(layer, section), new_shape

    def create_dense(self, section, in_features):
        sec = self.parser[section]
        units = sec.get('units')
        if not units:
            assert len(self.output_units) > self.output_count, 'Output units given are less than dense layers required'
            units = self.output_units[self.output_count]
            self.output_count += 1
        layer = _Dense(in_features, int(units), sec.get('activation'), self.tensor_core_dense)
        return self._finish(layer, section), int(units)

    def build_model(self):
        sections = self.parser.sections()
        assert sections, f'Empty model configuration {self.cfg_file}'
        if self.seed is not None:
            self._generator = torch.Generator().manual_seed(int(self.seed))
        layers, plan, common, outputs, activations = [], [], [], [], []
        shape, common_shape = self.input_shape, None
        for section in sections:
            sec = self.parser[section]
            if section.startswith('convolutional'):
                made, shape = self.create_convolution(section, shape)
                plan.append(False)
            elif section.startswith('flatten'):
                made, shape = _Flatten(), (math.prod(shape),)
                plan.append(False)
            elif section.startswith('dense'):
                source = common_shape if common_shape is not None else shape
                # Keras' Dense contracts the last axis whatever the rank (no implicit flatten)
                made, units = self.create_dense(section, source[-1])
                shape = source[:-1] + (units,)
                plan.append(common_shape is not None)
            else:
                continue                                           # unknown sections are skipped, as in the reference
            layers.append(made)
            activations.append(sec.get('activation'))
            if sec.get('common'):
                common_shape = shape
                common.append(len(layers) - 1)
            if sec.get('output'):
                outputs.append(len(layers) - 1)
        self.output_count = 0
        assert outputs, f'{self.cfg_file}: no section is marked output=1'
        return CfgNetwork(layers, plan, common, outputs, self.input_shape, activations)

    def build_tensor_core_network(self):
        """The `.cfg` as the documented Conv2D network ON THE TCGEN05 KERNELS, forward and backward (`NatureCnnTc`): possible when
        the file describes exactly that network -- 84x84x4 frames, convolutions 32x8/4, 64x4/2, 64x3/1 (ReLU), flatten, one
        common Dense of 512 (ReLU), an actor head and a one-unit critic head without output activations
        (ppo/models/cnn-actor-critic.cfg:1-42).  The initial weights are the ones this reader's initialisers and seed produce
        for the file (the FC weight re-ordered from Keras' (h, w, c) flatten to the module's (c, h, w))."""
        from .tc_cnn import NatureCnnTc
        assert self.conv_dims == 2, 'the tensor-core network is the Conv2D reading of the file: conv_dims=2'
        ref = self.build_model()
        layers = list(ref.layers)
        convs = [l for l in layers if isinstance(l, _ConvChannelsLast)]
        dense = [l for l in layers if isinstance(l, _Dense)]
        spec = [(c.conv.in_channels, c.conv.out_channels, c.conv.kernel_size[0], c.conv.stride[0]) for c in convs]
        ok = (self.input_shape == (84, 84, 4) and spec == [(4, 32, 8, 4), (32, 64, 4, 2), (64, 64, 3, 1)]
              and all(isinstance(c.act, torch.nn.ReLU) for c in convs) and len(dense) == 3
              and (dense[0].weight.shape[0], dense[0].weight.shape[1]) == (512, 3136) and isinstance(dense[0].act, torch.nn.ReLU)
              and dense[1].act is None and dense[2].act is None and dense[2].weight.shape[0] == 1 and not ref.output_is_softmax
              and len(ref.output_indices) == 2 and dense[1].weight.shape[0] <= 7)
        if not ok:
            raise ValueError(f'{self.cfg_file}: not the documented 84x84x4 Nature CNN actor-critic (32x8/4, 64x4/2, 64x3/1, Dense 512, actor <= 7 '
                             f'units + critic heads without output activations) -- the tensor-core network implements exactly that one')
        net = NatureCnnTc(4, dense[1].weight.shape[0])
        mine = [m for m in net.trunk if isinstance(m, torch.nn.Conv2d)]
        fc = [m for m in net.trunk if isinstance(m, torch.nn.Linear)][0]
        with torch.no_grad():
            for dst, src in zip(mine, convs):
                dst.weight.copy_(src.weight)
                dst.bias.copy_(src.bias)
            fc.weight.copy_(dense[0].weight.view(512, 49, 64).permute(0, 2, 1).reshape(512, 3136))
            fc.bias.copy_(dense[0].bias)
            net.actor.weight.copy_(dense[1].weight), net.actor.bias.copy_(dense[1].bias)
            net.critic.weight.copy_(dense[2].weight), net.critic.bias.copy_(dense[2].bias)
        net.output_is_softmax = False
        return net

    def build_adapter(self, device='cuda:0', tensor_core_network=False, **adapter_kwargs):
        """Network on `device` inside the `TorchModel` adapter the agents drive, optimiser keywords applied."""
        from .models import TorchModel
        opt = dict(self.optimizer or {})
        kw = dict(lr=opt.get('learning_rate', 7e-4), beta1=opt.get('beta_1', 0.9), beta2=opt.get('beta_2', 0.999),
                  epsilon=opt.get('epsilon', 1e-7))
        kw.update(adapter_kwargs)
        net = (self.build_tensor_core_network() if tensor_core_network else self.build_model()).to(device)
        return TorchModel(net, output_is_softmax=net.output_is_softmax, **kw)
