#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own code (read-only, /root/reference).

TensorFlow, tensorflow-probability, gym, optuna, termcolor and matplotlib are not installable in
this image, so `import xagents` cannot work as is.  This script installs import-time stand-ins:

* permissive stub modules for everything the reference imports but the hot path never calls;
* a NumPy-backed shim of exactly the TensorFlow / tfp ops the hot path calls
  (range, random.shuffle, gather, reduce_mean, math.reduce_std, clip_by_value, square, maximum,
  exp, squeeze, cast, minimum, reshape, split, stack, numpy_function, function, GradientTape, clip_by_global_norm;
  Categorical.log_prob / entropy / sample) following their published definitions, fp32;
* fake gym environments that replay a seeded synthetic stream, and a tiny NumPy "model".

It then constructs the genuine `xagents.PPO` / `xagents.A2C` objects and calls their genuine
`train_step()` / `calculate_returns()` / `concat_step_batches()` / `get_mini_batches()` /
`run_ppo_epochs()` / `update_gradients()`; thin recorders wrapped around the bound methods save
inputs and outputs.  Nothing from the reference is copied into this repository -- only the arrays
it produced.  The fixtures travel to the GPU box; /root/reference does not.

Usage:  python tests/golden/make_golden.py   (needs /root/reference; writes next to this file)
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get('XAGENTS_REFERENCE', '/root/reference')
STUB_ROOTS = ('tensorflow', 'tensorflow_probability', 'gym', 'optuna', 'termcolor', 'matplotlib')
F32 = np.float32


# ------------------------------------------------------------------ permissive stubs
class _StubMeta(type):
    def __getattr__(cls, name):
        if name.startswith('__'):
            raise AttributeError(name)
        value = _make_stub(f'{cls.__name__}.{name}')
        type.__setattr__(cls, name, value)     # stable identity: `tf.keras.activations.softmax in activations` must work
        return value

    def __call__(cls, *args, **kwargs):
        # decorator use (tf.function): hand the function back untouched
        if len(args) == 1 and not kwargs and isinstance(args[0], types.FunctionType):
            return args[0]
        return super().__call__()


def _make_stub(name):
    return _StubMeta(name, (), {'__init__': lambda self, *a, **k: None,
                                '__getattr__': lambda self, n: _make_stub(n),
                                '__call__': lambda self, *a, **k: None})


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        value = _make_stub(name)
        setattr(self, name, value)
        return value


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split('.')[0] in STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        _populate(module)


# ------------------------------------------------------------------ numpy-backed TF / tfp shim
RECORD = {'shuffles': [], 'means': [], 'losses': []}
_SHUFFLE_RNG = np.random.default_rng(4321)
_SAMPLE_RNG = np.random.default_rng(99)


def _f32(x):
    return np.asarray(x, F32)


class _Tensor(np.ndarray):
    """ndarray that answers .numpy() like an eager tensor."""

    def numpy(self):
        return np.asarray(self)


def _shuffle(indices):
    out = _SHUFFLE_RNG.permutation(np.asarray(indices)).astype(np.int32)
    RECORD['shuffles'].append(out.copy())
    return out


def _reduce_mean(x):
    m = np.asarray(x).mean(dtype=F32)
    RECORD['means'].append(F32(m))
    return m


def _reduce_std(x):
    x = np.asarray(x)
    return np.sqrt(np.square(x - x.mean(dtype=F32)).mean(dtype=F32))


class _Tape:
    def __init__(self, *args, **kwargs):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def gradient(self, loss, variables):
        RECORD['losses'].append(F32(loss))
        return []


def _log_softmax(x):
    z = x - x.max(axis=-1, keepdims=True)
    return z - np.log(np.exp(z).sum(axis=-1, keepdims=True))


class Categorical:
    """tfp.distributions.Categorical (0.15.0) restated: see SURVEY.md appendix A."""

    def __init__(self, logits=None, probs=None):
        raw = _f32(logits) if logits is not None else np.log(_f32(probs))
        self.lsm = _log_softmax(raw)

    def log_prob(self, actions):
        a = np.asarray(actions).astype(np.int64).reshape(-1, 1)
        return np.take_along_axis(self.lsm, a, axis=-1)[:, 0]

    def entropy(self):
        return -(np.exp(self.lsm) * self.lsm).sum(axis=-1, dtype=F32)

    def kl_divergence(self, other):
        """KL(self || other) = sum softmax(a) * (log_softmax(a) - log_softmax(b)) (tfp's Categorical-Categorical KL)."""
        return (np.exp(self.lsm) * (self.lsm - other.lsm)).sum(axis=-1, dtype=F32)

    def sample(self, seed=None):
        g = _SAMPLE_RNG.gumbel(size=self.lsm.shape)
        return np.argmax(self.lsm + g, axis=-1).astype(np.int32)


class MultivariateNormalDiag:
    def __init__(self, loc):
        self.loc = _f32(loc)

    def log_prob(self, actions):
        k = self.loc.shape[-1]
        d = _f32(actions).reshape(self.loc.shape) - self.loc
        return F32(-0.5) * np.square(d).sum(-1, dtype=F32) - F32(0.5 * k) * F32(np.log(2 * np.pi))

    def entropy(self):
        k = self.loc.shape[-1]
        return np.full(self.loc.shape[0], F32(0.5 * k) * (F32(1) + F32(np.log(2 * np.pi))), F32)

    def sample(self, seed=None):
        return (self.loc + _SAMPLE_RNG.standard_normal(self.loc.shape)).astype(F32)


class Discrete:
    def __init__(self, n):
        self.n = n

    def seed(self, s):
        pass


class Box:
    def __init__(self, shape):
        self.shape = shape

    def seed(self, s):
        pass


def _populate(module):
    name = module.__name__
    if name == 'tensorflow':
        module.function = lambda fn: fn
        module.float32 = 'float32'
        module.numpy_function = lambda fn, args, dtypes: fn(*args)
        module.range = lambda n: np.arange(n, dtype=np.int32)
        module.gather = lambda item, idx: np.take(np.asarray(item), np.asarray(idx), axis=0)
        module.reduce_mean = _reduce_mean
        module.clip_by_value = lambda x, lo, hi: np.clip(x, F32(lo), F32(hi))
        module.square = np.square
        module.maximum = np.maximum
        module.exp = np.exp
        module.squeeze = lambda x, axis=None: np.squeeze(np.asarray(x), axis).view(_Tensor)
        module.minimum = np.minimum
        module.reshape = lambda t, shape: np.reshape(np.asarray(t), shape)
        module.split = lambda t, n, axis=0: np.split(np.asarray(t), n, axis)
        module.stack = lambda items, axis=0: np.stack([np.asarray(i) for i in items], axis)
        module.cast = lambda x, dtype: np.asarray(x).astype(np.int64) if dtype == 'int64' else _f32(x)
        module.int64, module.int32, module.newaxis = 'int64', 'int32', None
        module.reduce_sum = lambda x, axis=None: np.asarray(x).sum(axis=axis, dtype=F32)
        module.stop_gradient = lambda x: x
        module.concat = lambda items, axis=0: np.concatenate([np.asarray(i) for i in items], axis)
        module.gather_nd = lambda params, idx: np.asarray(params)[tuple(np.asarray(idx).T)]
        module.tensordot = lambda a, b, axes: np.tensordot(a, b, axes)
        module.GradientTape = _Tape
        module.clip_by_global_norm = lambda grads, norm: (grads, F32(0))
    elif name == 'tensorflow.random':
        module.shuffle = _shuffle
        module.set_seed = lambda s: None
    elif name == 'tensorflow.math':
        module.reduce_std = _reduce_std
        module.log = np.log
        module.sqrt = np.sqrt
    elif name == 'tensorflow_probability.python.distributions':
        module.Categorical = Categorical
        module.MultivariateNormalDiag = MultivariateNormalDiag
    elif name in ('gym.spaces', 'gym.spaces.discrete', 'gym.spaces.box'):
        module.Discrete = Discrete
        module.Box = Box


def _import_reference():
    sys.meta_path.insert(0, _StubFinder())
    sys.path.insert(0, REFERENCE)
    import tensorflow  # noqa: F401  (the stub)
    import tensorflow.math  # noqa: F401
    import tensorflow.random  # noqa: F401
    import xagents
    assert os.path.realpath(xagents.__file__).startswith(os.path.realpath(REFERENCE))
    return xagents


# ------------------------------------------------------------------ fake envs and model
class ReplayEnv:
    """Replays a fixed stream: step k -> (obs[k+1], reward[k], done[k]); reset -> fresh obs."""

    def __init__(self, obs, rewards, dones, resets, action_space):
        self.obs, self.rewards, self.dones, self.resets = obs, rewards, dones, resets
        self.k = 0
        self.n_resets = 0
        self.action_space = action_space
        self.observation_space = types.SimpleNamespace(shape=obs.shape[1:])
        self.spec = types.SimpleNamespace(id='Replay-v0')

    def reset(self):
        out = self.resets[self.n_resets % len(self.resets)]
        self.n_resets += 1
        return out

    def step(self, action):
        k = self.k
        self.k += 1
        return self.obs[k + 1], float(self.rewards[k]), bool(self.dones[k]), {}

    def seed(self, s):
        pass


class TinyModel:
    """[actor_out, critic_out] = fixed fp32 linear maps of the first features of the input."""

    def __init__(self, in_features, n_actions, rng, softmax=False):
        self.wa = (rng.standard_normal((in_features, n_actions)) * 0.7).astype(F32)
        self.wc = (rng.standard_normal((in_features, 1)) * 0.7).astype(F32)
        self.drift = F32(0.0)          # set > 0 to emulate "weights moved since the rollout"
        self.softmax = softmax
        act = sys.modules['tensorflow'].keras.activations.softmax if softmax else None
        self.layers = [types.SimpleNamespace(activation=None), types.SimpleNamespace(activation=act),
                       types.SimpleNamespace(activation=None)]
        self.trainable_variables = []
        self.optimizer = types.SimpleNamespace(apply_gradients=lambda pairs: None)
        self.calls = []

    def __call__(self, inputs, training=True):
        x = _f32(inputs).reshape(len(inputs), -1)[:, :self.wa.shape[0]]
        actor = x @ (self.wa * (F32(1) + self.drift)) + self.drift
        critic = x @ (self.wc * (F32(1) - self.drift)) - self.drift
        if self.softmax:
            actor = np.exp(_log_softmax(actor))
        self.calls.append((actor.copy(), critic.copy()))
        return [actor, critic]


def _streams(rng, n_steps, n_envs, obs_shape, image, p_done):
    if image:
        obs = rng.integers(0, 256, size=(n_envs, n_steps + 2) + obs_shape, dtype=np.uint8)
        resets = rng.integers(0, 256, size=(n_envs, 4) + obs_shape, dtype=np.uint8)
    else:
        obs = rng.standard_normal((n_envs, n_steps + 2) + obs_shape).astype(F32)
        resets = rng.standard_normal((n_envs, 4) + obs_shape).astype(F32)
    rewards = rng.standard_normal((n_envs, n_steps + 1)).astype(F32)
    dones = rng.random((n_envs, n_steps + 1)) < p_done
    return obs, rewards, dones, resets


def _wrap(agent, name, sink):
    inner = getattr(agent, name)

    def recorder(*args, **kwargs):
        out = inner(*args, **kwargs)
        sink.append((args, out))
        return out

    setattr(agent, name, recorder)


def _build(xagents, kind, seed, n_steps, n_envs, obs_shape, image, n_actions, p_done, softmax=False, box=False, **kw):
    rng = np.random.default_rng(seed)
    obs, rewards, dones, resets = _streams(rng, n_steps, n_envs, obs_shape, image, p_done)
    space = Box((n_actions,)) if box else Discrete(n_actions)      # Box: n_actions = action dimension k
    envs = [ReplayEnv(obs[i], rewards[i], dones[i], resets[i], space) for i in range(n_envs)]
    model = TinyModel(min(int(np.prod(obs_shape)), 24), n_actions, rng, softmax)
    cls = xagents.PPO if kind == 'ppo' else xagents.A2C
    agent = cls(envs, model, n_steps=n_steps, quiet=True, **kw)
    return agent, model


def ppo_case(xagents, tag, seed, n_steps, n_envs, obs_shape, image, n_actions, p_done, drift=0.05, **kw):
    for k in RECORD:
        RECORD[k].clear()
    actor_kind = 'normal' if kw.get('box') else ('probs' if kw.get('softmax') else 'logits')
    agent, model = _build(xagents, 'ppo', seed, n_steps, n_envs, obs_shape, image, n_actions, p_done, **kw)
    rec = {k: [] for k in ('calculate_returns', 'get_batch', 'get_mini_batches', 'update_gradients')}
    for name in rec:
        _wrap(agent, name, rec[name])
    # weights "move" between the rollout and the updates so that ratio != 1 and clips engage
    inner_epochs = agent.run_ppo_epochs

    def run_epochs(*batch):
        model.drift = F32(drift)
        model.calls.clear()
        return inner_epochs(*batch)

    agent.run_ppo_epochs = run_epochs
    agent.train_step()                                     # the reference's own PPO.train_step
    update_calls = list(model.calls)
    (rewards, dones, values), returns = rec['calculate_returns'][0]
    _, batch = rec['get_batch'][0]
    _, minibatches = rec['get_mini_batches'][0]
    n_mb = len(minibatches)
    next_values = None
    # bootstrap value = critic of the model call made inside calculate_returns: recompute
    model.drift = F32(0)
    next_values = np.squeeze(model(_f32(agent.get_states()) / F32(255.0) if image else _f32(agent.get_states()))[1])
    out = dict(
        actor_kind=actor_kind,
        n_steps=n_steps, n_envs=n_envs, gamma=agent.gamma, lam=agent.lam, clip_norm=agent.clip_norm,
        entropy_coef=agent.entropy_coef, value_loss_coef=agent.value_loss_coef,
        advantage_epsilon=agent.advantage_epsilon, mini_batch_size=agent.mini_batch_size,
        ppo_epochs=agent.ppo_epochs, steps_after=agent.steps,
        rewards=rewards, dones=dones, values=values, next_values=np.atleast_1d(next_values), returns=returns,
        flat_states=batch[0], flat_actions=batch[1], flat_returns=batch[2], flat_values=batch[3],
        flat_log_probs=batch[4],
        shuffles=np.stack(RECORD['shuffles']),
        losses=np.asarray(RECORD['losses'], F32),
        means=np.asarray(RECORD['means'], F32).reshape(n_mb, 4),   # per update: adv mean, entropy, max-vl mean, pg
    )
    assert len(rec['update_gradients']) == n_mb == len(update_calls)
    for i, ((args, _), (actor, critic)) in enumerate(zip(rec['update_gradients'], update_calls)):
        states, actions, old_values, rets, old_logp, adv = args
        mb = minibatches[i]
        out[f'mb{i}_states'] = mb[0]
        out[f'mb{i}_actions'] = np.asarray(actions)
        out[f'mb{i}_old_values'] = np.asarray(old_values)
        out[f'mb{i}_returns'] = np.asarray(rets)
        out[f'mb{i}_old_log_probs'] = np.asarray(old_logp)
        out[f'mb{i}_advantages'] = np.asarray(adv)
        out[f'mb{i}_actor'] = actor
        out[f'mb{i}_critic'] = np.squeeze(critic)
    # time-major rollout as the reference assembled it (lists of per-step arrays -> asarray fp32)
    states_tm = batch[0].reshape((n_envs, n_steps) + batch[0].shape[1:]).swapaxes(0, 1)
    out['states_time_major'] = np.ascontiguousarray(states_tm)
    np.savez_compressed(os.path.join(HERE, f'{tag}.npz'), **out)
    print(f'{tag}: T={n_steps} E={n_envs} minibatches={n_mb} loss[0]={out["losses"][0]:.6f}')


def a2c_case(xagents, tag, seed, n_steps, n_envs, obs_shape, image, n_actions, p_done, **kw):
    for k in RECORD:
        RECORD[k].clear()
    actor_kind = 'normal' if kw.get('box') else ('probs' if kw.get('softmax') else 'logits')
    agent, model = _build(xagents, 'a2c', seed, n_steps, n_envs, obs_shape, image, n_actions, p_done, **kw)
    rec = {k: [] for k in ('calculate_returns', 'np_train_step')}
    for name in rec:
        _wrap(agent, name, rec[name])
    inner = agent.np_train_step

    def after_batch():
        out = inner()
        model.drift = F32(0.05)
        model.calls.clear()
        return out

    agent.np_train_step = after_batch
    agent.train_step()                                     # the reference's own A2C.train_step
    (rewards, dones), returns = rec['calculate_returns'][0]
    _, (states, flat_returns, actions, old_values) = rec['np_train_step'][0]
    actor, critic = model.calls[0]
    model.drift = F32(0)
    x = _f32(agent.get_states()) / F32(255.0) if image else _f32(agent.get_states())
    next_values = np.squeeze(model(x)[1])
    np.savez_compressed(
        os.path.join(HERE, f'{tag}.npz'),
        actor_kind=actor_kind, n_steps=n_steps, n_envs=n_envs, gamma=agent.gamma, entropy_coef=agent.entropy_coef,
        value_loss_coef=agent.value_loss_coef, steps_after=agent.steps,
        rewards=rewards, dones=dones, next_values=np.atleast_1d(next_values), returns=returns,
        flat_states=states, flat_returns=flat_returns, flat_actions=actions, flat_values=old_values,
        actor=actor, critic=np.squeeze(critic), loss=np.asarray(RECORD['losses'], F32),
        means=np.asarray(RECORD['means'], F32))            # entropy, adv*logp mean, value mse
    print(f'{tag}: T={n_steps} E={n_envs} loss={RECORD["losses"][0]:.6f}')


def kat_case(xagents):
    """SURVEY.md 8c KAT-1/KAT-2 inputs through the reference's calculate_returns + flatten."""
    T, E = 4, 3
    rewards = _f32([[1, 0, -1], [0, 2, .5], [1, 1, 1], [.5, 0, -2]])
    values = _f32([[.5, .1, -.3], [.2, .4, 0], [-.1, .3, .7], [.9, -.5, .2]])
    dones = _f32([[0, 0, 0], [0, 1, 0], [0, 0, 0], [1, 0, 0], [0, 0, 1]])
    next_values = _f32([.3, -.2, .6])

    def bare(cls):
        a = object.__new__(cls)
        a.n_steps, a.n_envs, a.gamma, a.lam, a.output_models = T, E, 0.99, 0.95, []
        a.get_states = lambda: None
        a.get_model_outputs = lambda *args, **kw: (
            None, None, types.SimpleNamespace(numpy=lambda: next_values.copy(), shape=(E,)), None, None)
        return a

    ppo_ret = xagents.PPO.calculate_returns(bare(xagents.PPO), rewards, dones, values)
    a2c = bare(xagents.A2C)
    a2c.get_model_outputs = lambda *args, **kw: (None, None, next_values.copy(), None, None)
    a2c_ret = xagents.A2C.calculate_returns(a2c, rewards, dones)
    flat = xagents.PPO.concat_step_batches(ppo_ret, values)
    np.savez_compressed(os.path.join(HERE, 'kat_returns.npz'), rewards=rewards, values=values, dones=dones,
                        next_values=next_values, gamma=0.99, lam=0.95, ppo_returns=ppo_ret,
                        a2c_returns=a2c_ret, flat_returns=flat[0], flat_values=flat[1])
    print('kat_returns:', ppo_ret[0], a2c_ret[0])


def acer_case(xagents):
    """ACER's Retrace returns (xagents/acer/agent.py:171-208) on flat env-major inputs, as ACER feeds them."""
    T, E = 11, 6
    rng = np.random.default_rng(31)
    flat = lambda *shape: rng.standard_normal(shape).astype(F32)
    rewards, q_sel = flat(E * T), flat(E * T)
    dones = (rng.random(E * T) < 0.15).astype(F32)
    values = flat(E * (T + 1))
    importance = np.exp(0.7 * rng.standard_normal(E * T)).astype(F32)
    a = object.__new__(xagents.ACER)
    a.n_steps, a.n_envs, a.gamma = T, E, 0.99
    out = xagents.ACER.calculate_returns(a, rewards, dones, values, q_sel, importance)
    np.savez_compressed(os.path.join(HERE, 'acer_retrace.npz'), n_steps=T, n_envs=E, gamma=0.99, rewards=rewards, dones=dones,
                        values=values, q_selected=q_sel, importance=importance, returns=np.asarray(out))
    print('acer_retrace:', np.asarray(out)[:4])


def _reset_rngs():
    """The shuffle / sampling streams are module-wide and consumed case after case; groups of cases added later start from
    the initial seeds again so that they do not depend on (or disturb) the cases generated before them."""
    global _SHUFFLE_RNG, _SAMPLE_RNG
    _SHUFFLE_RNG = np.random.default_rng(4321)
    _SAMPLE_RNG = np.random.default_rng(99)


def distribution_cases(xagents):
    """The other two branches of A2C.get_distribution (a2c/agent.py:50-63) through the reference's own train steps:
    MultivariateNormalDiag for Box action spaces (actions [N, k]) and Categorical(probs=) for softmax-output models."""
    _reset_rngs()
    ppo_case(xagents, 'ppo_box', 15, n_steps=10, n_envs=4, obs_shape=(6,), image=False, n_actions=3, p_done=0.15,
             mini_batches=4, ppo_epochs=2, box=True)
    ppo_case(xagents, 'ppo_softmax', 16, n_steps=12, n_envs=6, obs_shape=(5,), image=False, n_actions=5, p_done=0.1,
             mini_batches=3, ppo_epochs=2, softmax=True)
    a2c_case(xagents, 'a2c_box', 23, n_steps=6, n_envs=5, obs_shape=(4,), image=False, n_actions=2, p_done=0.2, box=True)
    a2c_case(xagents, 'a2c_softmax', 24, n_steps=7, n_envs=4, obs_shape=(4,), image=False, n_actions=3, p_done=0.2, softmax=True)


def _set_precision(dtype):
    """The shim computes in `F32`; numerical differentiation of the reference's own functions needs it in float64."""
    global F32
    F32 = dtype


class _KerasModelBase:
    """What `isinstance(models, tf.keras.models.Model)` (base.py:507) is pointed at while the update cases run."""


def _set_model_base(tf, value):
    holder = tf.keras.models
    (setattr if isinstance(holder, types.ModuleType) else type.__setattr__)(holder, 'Model', value)


def _install_model_base(tf):
    previous = tf.keras.models.Model
    _set_model_base(tf, _KerasModelBase)
    return previous


def _bare(cls, **attrs):
    a = object.__new__(cls)
    for k, v in attrs.items():
        setattr(a, k, v)
    return a


def acer_update_case(xagents, tag, trust_region, seed=51):
    """The reference's own ACER.update_gradients (acer/agent.py:262-339) end to end -- model outputs -> values -> clip_last_step ->
    gather_nd -> importance weights -> Retrace -> calculate_losses -> calculate_grads (trust-region projection) -- under the shim
    in float64.  The one thing a NumPy shim cannot give is the tape: `tape.gradient(loss, action_probs)` is answered by central
    differences of the REFERENCE'S calculate_losses (fp64, h = 1e-6), so the gradient too is the reference's, not a restatement;
    what the tape is then asked to push into the network (`output_grads`) is recorded."""
    import tensorflow as tf
    _set_precision(np.float64)
    previous_model_base = _install_model_base(tf)
    try:
        T, E, A, Fdim = 7, 6, 4, 5
        N = T * E
        rng = np.random.default_rng(seed)

        class Net(_KerasModelBase):
            def __init__(self, wa, wc):
                self.wa, self.wc = wa, wc
                self.trainable_variables = []
                self.optimizer = types.SimpleNamespace(apply_gradients=lambda pairs: None)

            def __call__(self, x, training=True):
                x = np.asarray(x, F32)
                return [np.exp(_log_softmax(x @ self.wa)), x @ self.wc]

        wa, wc, wavg = (rng.standard_normal((Fdim, A)) * 0.8 for _ in range(3))
        agent = _bare(xagents.ACER, n_steps=T, n_envs=E, n_actions=A, gamma=0.99, epsilon=1e-6, importance_c=10.0, delta=1,
                      trust_region=trust_region, entropy_coef=0.01, value_loss_coef=0.5, grad_norm=None, img_inputs=False,
                      output_is_softmax=True, seed=None, distribution_type=Categorical, model=Net(wa, wc), avg_model=Net(wavg, wc),
                      ema=types.SimpleNamespace(apply=lambda v: None), batch_indices=np.arange(N, dtype=np.int64)[:, None])
        agent.update_avg_weights = lambda: None
        states = rng.standard_normal((E * (T + 1), Fdim))                  # env-major, T+1 steps per env (acer/agent.py:146-162)
        rewards = rng.standard_normal(N)
        actions = rng.integers(0, A, N).astype(np.int32)
        dones = (rng.random(N) < 0.15).astype(np.float64)
        previous = np.exp(_log_softmax(rng.standard_normal((N, A))))
        rec = {'calculate_losses': [], 'calculate_grads': [], 'calculate_returns': []}
        inner_losses = agent.calculate_losses
        for name in rec:
            _wrap(agent, name, rec[name])
        pushed = {}

        class Tape(_Tape):
            def gradient(self, target, sources, output_gradients=None):
                if output_gradients is not None:                          # tape.gradient(action_probs, variables, output_grads)
                    pushed['output_grads'] = np.asarray(output_gradients).copy()
                    return []
                if isinstance(sources, np.ndarray) and sources.ndim == 2:  # tape.gradient(loss, action_probs)
                    (probs, values, returns, _, sel_imp, sel_q), _ = rec['calculate_losses'][-1]
                    return _fd_probs(inner_losses, probs, actions, values, returns, sel_imp, sel_q)
                return []

        tf.GradientTape, saved = Tape, tf.GradientTape
        try:
            agent.update_gradients(states, rewards, actions, dones, previous)
        finally:
            tf.GradientTape = saved
        (probs, values, returns, sel_probs, sel_imp, sel_q), losses = rec['calculate_losses'][0]
        (_, _, _, avg_probs), _ = rec['calculate_grads'][0]
        full_probs, full_q = agent.model(states)
        out = dict(n_steps=T, n_envs=E, n_actions=A, gamma=0.99, epsilon=1e-6, importance_c=10.0, delta=1, trust_region=trust_region,
                   entropy_coef=0.01, value_loss_coef=0.5, rewards=rewards, actions=actions, dones=dones, previous_action_probs=previous,
                   full_action_probs=full_probs, full_critic_logits=full_q, full_avg_action_probs=agent.avg_model(states)[0],
                   action_probs=np.asarray(probs), avg_action_probs=np.asarray(avg_probs), values=np.asarray(values),
                   returns=np.asarray(returns), selected_probs=np.asarray(sel_probs), selected_importance=np.asarray(sel_imp),
                   selected_critic_logits=np.asarray(sel_q))
        # gradients of the reference's loss w.r.t. the two model outputs, by central differences of the reference's function
        g = _fd_probs(inner_losses, probs, actions, values, returns, sel_imp, sel_q)
        dq = _fd_selected_q(inner_losses, probs, values, returns, sel_probs, sel_imp, sel_q, trust_region)
        d_critic = np.zeros((N, A))
        d_critic[np.arange(N), actions] = dq
        if trust_region:
            out.update(loss=float(losses[0]), value_loss=float(losses[1]), d_loss_d_action_probs=g,
                       output_grads=pushed['output_grads'], d_value_loss_d_critic_logits=d_critic)
        else:
            out.update(loss=float(losses), d_loss_d_action_probs=g, d_loss_d_critic_logits=d_critic)
        np.savez_compressed(os.path.join(HERE, f'{tag}.npz'), **out)
        print(f'{tag}: loss={out["loss"]:.6f} |g|max={np.abs(g).max():.4f}')
    finally:
        _set_precision(np.float32)
        _set_model_base(tf, previous_model_base)


def _fd_probs(calculate_losses, probs, actions, values, returns, sel_imp, sel_q, h=1e-6):
    """d loss / d action_probs[i, j] of the reference's calculate_losses; selected_probs follows action_probs (gather_nd)."""
    probs = np.asarray(probs, np.float64)
    n, a = probs.shape
    idx = np.arange(n)

    def f(p):
        out = calculate_losses(p, values, returns, p[idx, actions], sel_imp, sel_q)
        return float(out[0]) if isinstance(out, tuple) else float(out)
    g = np.zeros_like(probs)
    for i in range(n):
        for j in range(a):
            up, dn = probs.copy(), probs.copy()
            up[i, j] += h
            dn[i, j] -= h
            g[i, j] = (f(up) - f(dn)) / (2 * h)
    return g


def _fd_selected_q(calculate_losses, probs, values, returns, sel_probs, sel_imp, sel_q, trust_region, h=1e-6):
    sel_q = np.asarray(sel_q, np.float64)

    def f(q):
        out = calculate_losses(probs, values, returns, sel_probs, sel_imp, q)
        return float(out[1]) if trust_region else float(out)
    g = np.zeros_like(sel_q)
    for i in range(len(sel_q)):
        up, dn = sel_q.copy(), sel_q.copy()
        up[i] += h
        dn[i] -= h
        g[i] = (f(up) - f(dn)) / (2 * h)
    return g


def trpo_case(xagents):
    """The reference's own TRPO.calculate_losses / calculate_kl_divergence (trpo/agent.py:179-224) and update_critic_weights
    (:279-297) on linear NumPy networks: surrogate objective, mean KL(old || new), the critic's value loss per shuffled minibatch."""
    import tensorflow as tf
    _reset_rngs()
    for k in RECORD:
        RECORD[k].clear()
    T, E, A, Fdim = 9, 4, 5, 6
    N = T * E
    rng = np.random.default_rng(61)

    previous_model_base = _install_model_base(tf)

    class Linear(_KerasModelBase):
        def __init__(self, w):
            self.w = w
            self.trainable_variables = []
            self.optimizer = types.SimpleNamespace(apply_gradients=lambda pairs: None)

        def __call__(self, x, training=True):
            return _f32(x) @ self.w

    w_new, w_old = (rng.standard_normal((Fdim, A)) * 0.6).astype(F32), None
    w_old = (w_new + 0.05 * rng.standard_normal((Fdim, A))).astype(F32)
    w_critic = (rng.standard_normal((Fdim, 1)) * 0.6).astype(F32)
    actor, old_actor, critic = Linear(w_new), Linear(w_old), Linear(w_critic)
    agent = _bare(xagents.TRPO, n_steps=T, n_envs=E, entropy_coef=0.01, img_inputs=False, output_is_softmax=False, seed=None,
                  distribution_type=Categorical, actor=actor, old_actor=old_actor, critic=critic, output_models=[actor, critic],
                  critic_iterations=2, ppo_epochs=2, mini_batches=3, batch_size=N, mini_batch_size=N // 3)
    states = rng.standard_normal((N, Fdim)).astype(F32)
    actions = rng.integers(0, A, N).astype(F32)
    raw_adv = rng.standard_normal(N).astype(F32)
    advantages = ((raw_adv - raw_adv.mean(dtype=F32)) / _reduce_std(raw_adv)).astype(F32)        # trpo/agent.py:316-319
    surrogate, kl = xagents.TRPO.calculate_losses(agent, states, actions, advantages)
    returns = rng.standard_normal(N).astype(F32)
    RECORD['losses'].clear()
    RECORD['shuffles'].clear()
    xagents.TRPO.update_critic_weights(agent, states, returns)
    np.savez_compressed(os.path.join(HERE, 'trpo_losses.npz'), n_steps=T, n_envs=E, n_actions=A, entropy_coef=0.01, w_actor=w_new,
                        w_old_actor=w_old, w_critic=w_critic, states=states, actions=actions, raw_advantages=raw_adv,
                        advantages=advantages, surrogate_loss=F32(surrogate), kl_divergence=F32(kl), returns=returns,
                        critic_iterations=2, ppo_epochs=2, mini_batches=3, critic_value_losses=np.asarray(RECORD['losses'], F32),
                        critic_shuffles=np.stack(RECORD['shuffles']))
    _set_model_base(tf, previous_model_base)
    print(f'trpo_losses: surrogate={float(surrogate):.6f} kl={float(kl):.6e} value losses={len(RECORD["losses"])}')


class _PerturbableModel(TinyModel):
    """TinyModel whose outputs can be nudged entry by entry: central differences of the reference's own loss w.r.t. the
    model outputs stand in for the tape (the shim has none)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.delta_actor = self.delta_critic = None

    def __call__(self, inputs, training=True):
        actor, critic = super().__call__(inputs, training)
        if self.delta_actor is not None:
            actor = actor + self.delta_actor
        if self.delta_critic is not None:
            critic = critic + self.delta_critic.reshape(critic.shape)
        return [actor, critic]


def loss_gradient_case(xagents, tag, kind, seed, softmax=False, box=False, n_actions=4, h=1e-6):
    """d loss / d(actor_output, critic_output) of the REFERENCE'S OWN PPO.update_gradients (ppo/agent.py:96-134) and
    A2C.train_step loss (a2c/agent.py:202-214): the reference code runs under the shim in float64 and its recorded loss is
    differenced centrally (h = 1e-6) over every entry of the model outputs.  This pins the gradients the CUDA loss kernels
    return (ties of maximum / clip_by_value are a set of measure zero for these inputs)."""
    _set_precision(np.float64)
    try:
        T, E, Fdim = 6, 4, 5
        N = T * E
        rng = np.random.default_rng(seed)
        obs, rewards, dones, resets = _streams(rng, T, E, (Fdim,), False, 0.1)
        space = Box((n_actions,)) if box else Discrete(n_actions)
        envs = [ReplayEnv(obs[i], rewards[i], dones[i], resets[i], space) for i in range(E)]
        model = _PerturbableModel(Fdim, n_actions, rng, softmax)
        model.wa, model.wc = model.wa.astype(np.float64), model.wc.astype(np.float64)
        cls = xagents.PPO if kind == 'ppo' else xagents.A2C
        kw = dict(mini_batches=1, ppo_epochs=1) if kind == 'ppo' else {}
        agent = cls(envs, model, n_steps=T, quiet=True, **kw)
        states = rng.standard_normal((N, Fdim))
        base_actor, base_critic = TinyModel.__call__(model, states)
        if box:
            actions = base_actor + rng.standard_normal((N, n_actions))
        else:
            actions = rng.integers(0, n_actions, N).astype(np.float64)
        old_values = np.squeeze(base_critic) + 0.08 * rng.standard_normal(N)        # both sides of the value clip (0.1)
        returns = rng.standard_normal(N)
        dist = agent.get_distribution(base_actor)
        old_log_probs = np.asarray(dist.log_prob(actions)) + 0.08 * rng.standard_normal(N)   # ratios on both sides of the clip band
        advantages = rng.standard_normal(N)

        def loss_of():
            RECORD['losses'].clear()
            if kind == 'ppo':
                agent.update_gradients(states, actions, old_values, returns, old_log_probs, advantages)
            else:                                                  # A2C.train_step's loss block, fed the same batch
                agent.np_train_step = lambda: (states, returns, actions, old_values)
                agent.train_step()
            return float(RECORD['losses'][-1])

        loss = loss_of()
        d_actor, d_critic = np.zeros((N, n_actions)), np.zeros(N)
        for i in range(N):
            for j in range(n_actions):
                model.delta_actor = np.zeros((N, n_actions))
                model.delta_actor[i, j] = h
                up = loss_of()
                model.delta_actor[i, j] = -h
                d_actor[i, j] = (up - loss_of()) / (2 * h)
            model.delta_actor = None
            model.delta_critic = np.zeros(N)
            model.delta_critic[i] = h
            up = loss_of()
            model.delta_critic[i] = -h
            d_critic[i] = (up - loss_of()) / (2 * h)
            model.delta_critic = None
        np.savez_compressed(os.path.join(HERE, f'{tag}.npz'), kind=kind, actor_kind='normal' if box else ('probs' if softmax else 'logits'),
                            n_actions=n_actions, clip_norm=getattr(agent, 'clip_norm', 0.0), entropy_coef=agent.entropy_coef,
                            value_loss_coef=agent.value_loss_coef, actor_output=base_actor, critic_output=np.squeeze(base_critic),
                            actions=actions, old_values=old_values, returns=returns, old_log_probs=old_log_probs, advantages=advantages,
                            loss=loss, d_actor=d_actor, d_critic=d_critic)
        print(f'{tag}: loss={loss:.6f} |d_actor|max={np.abs(d_actor).max():.5f} |d_critic|max={np.abs(d_critic).max():.5f}')
    finally:
        _set_precision(np.float32)


def gradient_cases(xagents):
    _reset_rngs()
    loss_gradient_case(xagents, 'grad_ppo_logits', 'ppo', 71)
    loss_gradient_case(xagents, 'grad_ppo_probs', 'ppo', 72, softmax=True)
    loss_gradient_case(xagents, 'grad_ppo_normal', 'ppo', 73, box=True, n_actions=3)
    loss_gradient_case(xagents, 'grad_a2c_logits', 'a2c', 74)
    loss_gradient_case(xagents, 'grad_a2c_probs', 'a2c', 75, softmax=True)
    loss_gradient_case(xagents, 'grad_a2c_normal', 'a2c', 76, box=True, n_actions=3)


def update_cases(xagents):
    acer_update_case(xagents, 'acer_update_trust_region', True)
    acer_update_case(xagents, 'acer_update_plain', False)
    trpo_case(xagents)



def main():
    xagents = _import_reference()
    if '--acer-only' in sys.argv:
        return acer_case(xagents)
    if '--gradients-only' in sys.argv:                             # round 2: leaves the earlier fixtures untouched
        return gradient_cases(xagents)
    if '--updates-only' in sys.argv:                               # round 2: leaves the earlier fixtures untouched
        return update_cases(xagents)
    if '--distributions-only' in sys.argv:                         # added later: leaves the earlier fixtures untouched
        return distribution_cases(xagents)
    acer_case(xagents)
    kat_case(xagents)
    # PPO, image observations (uint8-valued, 8x8x4), 6 actions: Atari-shaped in miniature
    ppo_case(xagents, 'ppo_image', 11, n_steps=16, n_envs=8, obs_shape=(8, 8, 4), image=True,
             n_actions=6, p_done=0.08, mini_batches=4, ppo_epochs=4)
    # PPO, CartPole-shaped vectors (C1 in miniature: E=16, T=128, A=2)
    ppo_case(xagents, 'ppo_cartpole', 12, n_steps=128, n_envs=16, obs_shape=(4,), image=False,
             n_actions=2, p_done=0.02, mini_batches=4, ppo_epochs=4)
    # PPO with N % mini_batches != 0 -> trailing short minibatch; single env; harsher clip
    ppo_case(xagents, 'ppo_ragged', 13, n_steps=7, n_envs=3, obs_shape=(5,), image=False,
             n_actions=3, p_done=0.3, mini_batches=4, ppo_epochs=2, clip_norm=0.02, drift=0.2)
    ppo_case(xagents, 'ppo_single_env', 14, n_steps=9, n_envs=1, obs_shape=(6, 6, 1), image=True,
             n_actions=4, p_done=0.2, mini_batches=3, ppo_epochs=2)
    # A2C C2-shaped (E=16, T=5) with small frames, and a vector case
    a2c_case(xagents, 'a2c_image', 21, n_steps=5, n_envs=16, obs_shape=(8, 8, 4), image=True,
             n_actions=6, p_done=0.1)
    a2c_case(xagents, 'a2c_vector', 22, n_steps=12, n_envs=5, obs_shape=(4,), image=False,
             n_actions=2, p_done=0.15)
    distribution_cases(xagents)
    update_cases(xagents)
    gradient_cases(xagents)


if __name__ == '__main__':
    main()
