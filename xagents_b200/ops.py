"""Tensor-level entry points of the hot path: DLPack in, kernels through the C ABI, tensors out.

Arguments may be torch CUDA tensors, any object with `__dlpack__`, or raw DLPack capsules (e.g.
`tf.experimental.dlpack.to_dlpack(t)`), all zero-copy.  Outputs are allocated with torch (device
memory and streams are torch's job here; the arithmetic is not) unless `out=` buffers are given.
Every call is asynchronous on `stream` (default: torch's current stream).  There is no CPU path.
"""
import ctypes

import torch

from . import _ffi
from ._dlpack import as_device_array
from ._ffi import (XA_ACTOR_LOGITS, XA_ACTOR_NORMAL, XA_ACTOR_PROBS, XA_GATHER_AUTO, XA_GATHER_BULK, XA_GATHER_VECTOR,
                   XA_MAX_FIELDS, XA_MOMENT_STRIDE, XA_SCAN_AUTO, XA_SCAN_CHUNKED, XA_SCAN_SEQUENTIAL)

__all__ = ['gae_returns', 'nstep_returns', 'retrace_returns', 'gather_rows', 'gather_fields', 'gather_minibatch', 'gather_rows_scaled',
           'policy_step', 'adv_moments', 'normalize_advantages', 'ppo_loss', 'a2c_loss', 'loss_workspace', 'grad_sumsq', 'grad_from_outputs', 'clip_adam', 'optim_workspace', 'gemm_bf16_tn', 'to_bf16', 'conv2d_nhwc_bf16', 'space_to_depth_u8_bf16', 'gather_s2d_u8_bf16', 'conv_wgrad_nhwc_bf16', 'gemm_bf16_atb', 'gather_cast_f32',
           'launch_count', 'reset_launch_count']

SCAN_MODES = {'auto': XA_SCAN_AUTO, 'sequential': XA_SCAN_SEQUENTIAL, 'chunked': XA_SCAN_CHUNKED}
GATHER_MODES = {'auto': XA_GATHER_AUTO, 'bulk': XA_GATHER_BULK, 'vector': XA_GATHER_VECTOR}
ACTOR_KINDS = {'logits': XA_ACTOR_LOGITS, 'probs': XA_ACTOR_PROBS, 'normal': XA_ACTOR_NORMAL}

_launches = 0          # kernels launched through this module (bench.py reports it as gpu_launches)


def launch_count():
    return _launches


def reset_launch_count():
    global _launches
    _launches = 0


def _count(n=1):
    global _launches
    _launches += n


def _stream(stream, device=None):
    if stream is None:
        return torch.cuda.current_stream(device).cuda_stream
    return getattr(stream, 'cuda_stream', stream)


def _call(arr, fn_name, *args):
    """Launch on the device that owns `arr` (a DeviceArray): the kernels go to the CUDA runtime's CURRENT device, so a tensor
    on cuda:1 in a process whose current device is cuda:0 needs the switch (and that device's current stream).  The last
    argument is the caller's `stream` (None = torch's current stream on that device)."""
    index = arr.device_id
    stream = args[-1]
    if index == torch.cuda.current_device():
        return _ffi.call(fn_name, *args[:-1], _stream(stream, index))
    with torch.cuda.device(index):
        return _ffi.call(fn_name, *args[:-1], _stream(stream, index))


def _dev(x, dtype=None):
    return as_device_array(x, dtype)


def _device_of(arr):
    return torch.device('cuda', arr.device_id)


def _ptr(x):
    return ctypes.c_void_p(x.ptr if x is not None else None)


def _tptr(t):
    return ctypes.c_void_p(t.data_ptr() if t is not None else None)


# ------------------------------------------------------------------------------------------ returns
def gae_returns(rewards, values, last_values, dones, gamma, lam, *, mode='auto', with_advantages=False, out=None,
                stream=None):
    """PPO.calculate_returns arithmetic (xagents/ppo/agent.py:80-94).

    rewards, values [T,E]; last_values [E] (bootstrap); dones [T+1,E]; fp32 on the device.
    Returns returns [T,E] (and advantages when asked).
    """
    r, v, lv, d = _dev(rewards, 'float32'), _dev(values, 'float32'), _dev(last_values, 'float32'), _dev(dones, 'float32')
    if len(r.shape) != 2:
        raise ValueError(f'rewards must be [n_steps, n_envs], got {r.shape}')
    T, E = r.shape
    if v.size != T * E or lv.size != E or d.size != (T + 1) * E:
        raise ValueError(f'shape mismatch: rewards {r.shape}, values {v.shape}, last_values {lv.shape}, dones {d.shape} '
                         f'(dones needs n_steps+1 rows)')
    dev = _device_of(r)
    ret = out if out is not None else torch.empty((T, E), dtype=torch.float32, device=dev)
    adv = torch.empty((T, E), dtype=torch.float32, device=dev) if with_advantages else None
    _call(r, 'xa_gae_f32', _ptr(r), _ptr(v), _ptr(lv), _ptr(d), _tptr(ret), _tptr(adv), T, E, float(gamma), float(lam),
              SCAN_MODES[mode], stream)
    _count()
    return (ret, adv) if with_advantages else ret


def nstep_returns(rewards, dones, last_values, gamma, *, mode='auto', out=None, stream=None):
    """A2C.calculate_returns arithmetic (xagents/a2c/agent.py:165-171)."""
    r, d, lv = _dev(rewards, 'float32'), _dev(dones, 'float32'), _dev(last_values, 'float32')
    if len(r.shape) != 2:
        raise ValueError(f'rewards must be [n_steps, n_envs], got {r.shape}')
    T, E = r.shape
    if lv.size != E or d.size != (T + 1) * E:
        raise ValueError(f'shape mismatch: rewards {r.shape}, last_values {lv.shape}, dones {d.shape}')
    ret = out if out is not None else torch.empty((T, E), dtype=torch.float32, device=_device_of(r))
    _call(r, 'xa_nstep_returns_f32', _ptr(r), _ptr(d), _ptr(lv), _tptr(ret), T, E, float(gamma), SCAN_MODES[mode],
              stream)
    _count()
    return ret


def retrace_returns(rewards, dones, values, last_values, q_selected, importance, gamma, *, out=None, stream=None):
    """ACER.calculate_returns (Retrace) arithmetic (xagents/acer/agent.py:198-208), time-major [T,E] fields."""
    arrs = [_dev(x, 'float32') for x in (rewards, dones, values, last_values, q_selected, importance)]
    T, E = arrs[0].shape
    if arrs[1].size != (T + 1) * E or arrs[3].size != E or any(a.size != T * E for a in (arrs[2], arrs[4], arrs[5])):
        raise ValueError('shape mismatch: [T,E] fields, dones [T+1,E], last_values [E]')
    ret = out if out is not None else torch.empty((T, E), dtype=torch.float32, device=_device_of(arrs[0]))
    _call(arrs[0], 'xa_retrace_f32', *[_ptr(a) for a in arrs], _tptr(ret), T, E, float(gamma), stream)
    _count()
    return ret


# ------------------------------------------------------------------------------------------ gathers
def _layout(time_major):
    if time_major is None:
        return 0, 0
    T, E = time_major
    return int(T), int(E)


def _torch_dtype(name):
    return getattr(torch, name)


def gather_rows(src, idx, *, time_major=None, mode='auto', out=None, stream=None):
    """tf.gather(src, idx) along axis 0 (xagents/ppo/agent.py:154), whole rows, bit-exact.

    src [rows, ...]; with time_major=(T, E) src is the time-major rollout [T, E, ...] and idx holds
    env-major flat sample ids (base.py:559-564), so no flatten copy is needed.
    """
    s, i = _dev(src), _dev(idx, 'int32')
    T, E = _layout(time_major)
    lead = 2 if T else 1
    n_rows = s.shape[0] * (s.shape[1] if T else 1)
    if T and (s.shape[0], s.shape[1]) != (T, E):
        raise ValueError(f'src {s.shape} is not time-major [{T},{E},...]')
    row_shape = s.shape[lead:]
    row_bytes = s.itemsize
    for k in row_shape:
        row_bytes *= k
    n = i.size
    dst = out if out is not None else torch.empty((n,) + tuple(row_shape), dtype=_torch_dtype(s.dtype), device=_device_of(s))
    _call(s, 'xa_gather_rows', _ptr(s), _ptr(i), _tptr(dst), n, row_bytes, n_rows, T, E, GATHER_MODES[mode], stream)
    _count()
    return dst


def _field_tables(fields, outs, n, dev):
    if len(fields) > XA_MAX_FIELDS:
        raise ValueError(f'at most {XA_MAX_FIELDS} fields per call')
    arrs = [_dev(f, 'float32') for f in fields]
    if outs is None:
        outs = [torch.empty((n,), dtype=torch.float32, device=dev) for _ in arrs]
    src = (ctypes.c_void_p * len(arrs))(*[a.ptr for a in arrs])
    dst = (ctypes.c_void_p * len(arrs))(*[o.data_ptr() for o in outs])
    return arrs, outs, src, dst


def gather_fields(fields, idx, *, time_major=None, out=None, stream=None):
    """The four scalar-per-sample tf.gather calls of ppo/agent.py:154 in one launch."""
    i = _dev(idx, 'int32')
    T, E = _layout(time_major)
    arrs, outs, src, dst = _field_tables(fields, out, i.size, _device_of(i))
    _call(i, 'xa_gather_fields_f32', src, dst, len(arrs), _ptr(i), i.size, T, E, stream)
    _count()
    return outs


def gather_minibatch(obs, fields, idx, *, time_major=None, mode='auto', out_obs=None, out_fields=None, stream=None):
    """Observation rows + scalar fields of one minibatch (or a whole epoch) in one call."""
    s, i = _dev(obs), _dev(idx, 'int32')
    T, E = _layout(time_major)
    lead = 2 if T else 1
    n_rows = s.shape[0] * (s.shape[1] if T else 1)
    row_shape = s.shape[lead:]
    row_bytes = s.itemsize
    for k in row_shape:
        row_bytes *= k
    n = i.size
    dev = _device_of(s)
    dst = out_obs if out_obs is not None else torch.empty((n,) + tuple(row_shape), dtype=_torch_dtype(s.dtype), device=dev)
    arrs, outs, fsrc, fdst = _field_tables(fields, out_fields, n, dev)
    _call(s, 'xa_gather_minibatch', _ptr(s), _tptr(dst), row_bytes, n_rows, fsrc, fdst, len(arrs), _ptr(i), n, T, E,
              GATHER_MODES[mode], stream)
    # bulk path: one kernel; vector path: rows kernel + fields kernel
    bulk = mode == 'bulk' or (mode == 'auto' and row_bytes >= 2048 and row_bytes % 16 == 0 and n * row_bytes >= (16 << 20))
    _count(1 if bulk else 1 + (len(arrs) > 0))
    return dst, outs


def gather_rows_scaled(src_u8, idx, *, time_major=None, out=None, stream=None):
    """gather + cast(uint8 -> fp32) / 255 (xagents/base.py:505-506) in one pass."""
    s, i = _dev(src_u8, 'uint8'), _dev(idx, 'int32')
    T, E = _layout(time_major)
    lead = 2 if T else 1
    n_rows = s.shape[0] * (s.shape[1] if T else 1)
    row_shape = s.shape[lead:]
    row_bytes = 1
    for k in row_shape:
        row_bytes *= k
    n = i.size
    dst = out if out is not None else torch.empty((n,) + tuple(row_shape), dtype=torch.float32, device=_device_of(s))
    _call(s, 'xa_gather_rows_u8_scaled_f32', _ptr(s), _ptr(i), _tptr(dst), n, row_bytes, n_rows, T, E, stream)
    _count()
    return dst


# ------------------------------------------------------------------------------------------ rollout-time policy
def policy_step(actor_out, *, actor_kind='logits', noise=None, seed=0, offset=0, out=None, stream=None, counter=None, advance=None):
    """sample() + log_prob() + entropy() of A2C.get_model_outputs (a2c/agent.py:80-94) for one env step.
    Returns (actions, log_probs, entropies); `out` may hold row views of the rollout buffers.  `counter`: a device int64 [1]
    holding the Philox offset base; the kernel draws at counter + `offset` and the counter is advanced behind it by `advance`
    (default 2A; 0 = not at all, the caller bumps it once per rollout with `bump_u64`) -- for captured graphs."""
    ao = _dev(actor_out, 'float32')
    n, A = ao.shape[0], ao.shape[-1]
    dev = _device_of(ao)
    actions, logp, ent = out if out is not None else (None, None, None)
    if actions is None:
        actions = torch.empty((n, A) if actor_kind == 'normal' else (n,), dtype=torch.float32, device=dev)
    if logp is None:
        logp = torch.empty((n,), dtype=torch.float32, device=dev)
    if ent is None:
        ent = torch.empty((n,), dtype=torch.float32, device=dev)
    if counter is not None:
        assert noise is None and counter.dtype == torch.int64 and counter.is_cuda
        advance = 2 * A if advance is None else int(advance)
        _call(ao, 'xa_policy_step_counter_f32', _ptr(ao), ACTOR_KINDS[actor_kind], int(seed), _tptr(counter), int(offset), advance,
              _tptr(actions), _tptr(logp), _tptr(ent), n, A, stream)
        _count(2 if advance else 1)
        return actions, logp, ent
    nz = _dev(noise, 'float32') if noise is not None else None
    _call(ao, 'xa_policy_step_f32', _ptr(ao), ACTOR_KINDS[actor_kind], _ptr(nz), int(seed), int(offset), _tptr(actions),
              _tptr(logp), _tptr(ent), n, A, stream)
    _count()
    return actions, logp, ent


def bump_u64(counter, delta, *, stream=None):
    """counter[0] += delta in stream order (a device int64 [1]: the Philox offsets a captured rollout reads)."""
    assert counter.dtype == torch.int64 and counter.is_cuda and counter.numel() == 1
    _call(_dev(counter), 'xa_bump_u64', _tptr(counter), int(delta), stream)
    _count()


def synth_env_step(pool, states, new_states, rewards, dones, *, p_reward, p_done, seed, counter=None, offset=0, episode_sums=None,
                   sums_log=None, stream=None):
    """One step of the device-resident synthetic Atari environments with the step_envs bookkeeping (base.py:408-426) in ONE launch:
    `new_states` (or None) <- the frames the step returns, `states` <- the frames the environments hold afterwards (post-reset where
    done), `rewards` / `dones` [n] fp32, `episode_sums` += rewards -> `sums_log`, then cut at dones.  All torch CUDA tensors, written in place."""
    assert pool.is_cuda and pool.dtype == torch.uint8 and pool.is_contiguous() and states.dtype == torch.uint8 and states.is_contiguous()
    n = states.shape[0]
    row_bytes = states[0].numel()
    assert pool[0].numel() == row_bytes and rewards.numel() == n and dones.numel() == n
    for t in (new_states, rewards, dones, episode_sums, sums_log):
        assert t is None or (t.is_cuda and t.is_contiguous() and t.device == states.device)
    assert rewards.dtype == dones.dtype == torch.float32
    _call(_dev(states), 'xa_synth_env_step_u8', _tptr(pool), pool.shape[0], row_bytes, _tptr(states), _tptr(new_states), _tptr(rewards), _tptr(dones),
          _tptr(episode_sums), _tptr(sums_log), n, float(p_reward), float(p_done), int(seed) & (2 ** 64 - 1), _tptr(counter), int(offset), stream)
    _count()


# ------------------------------------------------------------------------------------------ moments + losses
def adv_moments(returns, old_values, idx, mb_offsets, *, time_major=None, out=None, stream=None):
    """(count, mean, M2) of returns - old_values for each minibatch idx[mb_offsets[m]:mb_offsets[m+1]]
    (ppo/agent.py:180-183).  Returns float64 [n_minibatches, 4] on the device."""
    r, v = _dev(returns, 'float32'), _dev(old_values, 'float32')
    i = _dev(idx, 'int32') if idx is not None else None
    T, E = _layout(time_major)
    n_mb = len(mb_offsets) - 1
    offs = (ctypes.c_int64 * (n_mb + 1))(*[int(o) for o in mb_offsets])
    mom = out if out is not None else torch.empty((n_mb, XA_MOMENT_STRIDE), dtype=torch.float64, device=_device_of(r))
    _call(r, 'xa_adv_moments_f32', _ptr(r), _ptr(v), _ptr(i), offs, n_mb, T, E, _tptr(mom), stream)
    _count((n_mb + 63) // 64)
    return mom


def normalize_advantages(returns_mb, old_values_mb, advantage_epsilon=1e-8, *, idx=None, time_major=None, moments=None,
                         out=None, stream=None):
    """advantages_mb of PPO.run_ppo_epochs (ppo/agent.py:180-183): (adv - mean) / (population std + eps).
    `moments` ([4] or [parts,4] fp64) defaults to the moments of exactly these samples."""
    r, v = _dev(returns_mb, 'float32'), _dev(old_values_mb, 'float32')
    i = _dev(idx, 'int32') if idx is not None else None
    n = i.size if i is not None else r.size
    T, E = _layout(time_major)
    if moments is None:
        moments = adv_moments(returns_mb, old_values_mb, idx, [0, n], time_major=time_major, stream=stream)[0]
    mom = _dev(moments, 'float64')
    parts = 1 if len(mom.shape) == 1 else mom.shape[0]
    adv = out if out is not None else torch.empty((n,), dtype=torch.float32, device=_device_of(r))
    _call(r, 'xa_normalize_adv_f32', _ptr(r), _ptr(v), _ptr(i), n, T, E, _ptr(mom), parts, mom.shape[-1] if parts > 1 else 0,
              float(advantage_epsilon), _tptr(adv), stream)
    _count()
    return adv


def loss_workspace(n, device):
    """Zero-filled scratch for ppo_loss / a2c_loss over minibatches of up to n samples (reusable)."""
    nbytes = _ffi.lib().xa_loss_workspace_bytes(int(n))
    return torch.zeros(((nbytes + 15) // 16) * 2, dtype=torch.float64, device=device)


def _loss(fn, ppo, actor_out, values, actions, old_log_probs, old_values, returns, *, idx, time_major, advantages, moments,
          clip, entropy_coef, value_loss_coef, advantage_epsilon, actor_kind, need_grads, return_advantages, workspace,
          out, stream):
    ao, v = _dev(actor_out, 'float32'), _dev(values, 'float32')
    n = v.size
    n_actions = ao.size // n
    dev = _device_of(ao)
    args = _ffi.LossArgs()
    keep = [ao, v]

    def put(name, obj, dtype='float32'):
        if obj is None:
            setattr(args, name, None)
            return None
        arr = _dev(obj, dtype)
        keep.append(arr)
        setattr(args, name, arr.ptr)
        return arr

    args.actor_out, args.values = ao.ptr, v.ptr
    put('actions', actions)
    put('old_log_probs', old_log_probs)
    put('old_values', old_values)
    put('returns', returns)
    put('idx', idx, 'int32')
    args.n_steps, args.n_envs = _layout(time_major)
    put('advantages', advantages)
    if moments is not None:
        if isinstance(moments, tuple):          # (base tensor, offset, n_parts, part_stride), in doubles
            base, offset, parts, stride = moments
            mom = _dev(base, 'float64')
            keep.append(mom)
            args.moments = mom.ptr + 8 * int(offset)
            args.n_moment_parts, args.moment_part_stride = int(parts), int(stride)
        else:
            mom = put('moments', moments, 'float64')
            parts = 1 if len(mom.shape) == 1 else mom.shape[0]
            args.n_moment_parts, args.moment_part_stride = parts, (mom.shape[-1] if parts > 1 else 0)
    args.n, args.n_actions, args.actor_kind = n, n_actions, ACTOR_KINDS[actor_kind]
    args.clip, args.ent_coef, args.vf_coef, args.adv_eps = clip, entropy_coef, value_loss_coef, advantage_epsilon
    scalars, d_actor, d_values, adv_out = out if out is not None else (None, None, None, None)
    if scalars is None:
        scalars = torch.empty(4, dtype=torch.float32, device=dev)
    if need_grads and d_actor is None:
        d_actor = torch.empty((n, n_actions), dtype=torch.float32, device=dev)
    if need_grads and d_values is None:
        d_values = torch.empty((n,), dtype=torch.float32, device=dev)
    if return_advantages and adv_out is None:
        adv_out = torch.empty((n,), dtype=torch.float32, device=dev)
    if workspace is None:
        workspace = loss_workspace(n, dev)
    args.out_scalars, args.d_actor, args.d_values = scalars.data_ptr(), _tptr(d_actor), _tptr(d_values)
    args.advantages_out = _tptr(adv_out)
    args.workspace, args.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    _call(ao, fn, ctypes.byref(args), stream)
    _count()
    return scalars, d_actor, d_values, adv_out


def ppo_loss(actor_out, values, actions, old_log_probs, old_values, returns, *, idx=None, time_major=None,
             advantages=None, moments=None, clip_norm=0.1, entropy_coef=0.01, value_loss_coef=0.5,
             advantage_epsilon=1e-8, actor_kind='logits', need_grads=True, return_advantages=False, workspace=None,
             out=None, stream=None):
    """PPO.update_gradients loss + gradients w.r.t. the model outputs (ppo/agent.py:112-134), with the
    per-minibatch advantage normalisation of ppo/agent.py:180-183 folded in (from `moments`: one
    [4] row, [parts, 4] rows to combine, or a (base, offset, n_parts, part_stride) view in doubles
    into an all-gathered [ranks, minibatches, 4] buffer) unless normalised `advantages` are supplied.

    Returns (scalars[4] = loss, pg, value_loss, entropy; d_actor [n,A]; d_values [n]; advantages|None).
    """
    return _loss('xa_ppo_loss_f32', True, actor_out, values, actions, old_log_probs, old_values, returns, idx=idx,
                 time_major=time_major, advantages=advantages, moments=moments, clip=clip_norm, entropy_coef=entropy_coef,
                 value_loss_coef=value_loss_coef, advantage_epsilon=advantage_epsilon, actor_kind=actor_kind,
                 need_grads=need_grads, return_advantages=return_advantages, workspace=workspace, out=out, stream=stream)


def a2c_loss(actor_out, values, actions, old_values, returns, *, idx=None, time_major=None, entropy_coef=0.01,
             value_loss_coef=0.5, actor_kind='logits', need_grads=True, workspace=None, out=None, stream=None):
    """A2C.train_step loss + gradients (a2c/agent.py:202-215)."""
    return _loss('xa_a2c_loss_f32', False, actor_out, values, actions, None, old_values, returns, idx=idx,
                 time_major=time_major, advantages=None, moments=None, clip=0.0, entropy_coef=entropy_coef,
                 value_loss_coef=value_loss_coef, advantage_epsilon=0.0, actor_kind=actor_kind, need_grads=need_grads,
                 return_advantages=False, workspace=workspace, out=out, stream=stream)[:3]


# ------------------------------------------------------------------------------------------ optimiser
def optim_workspace(device):
    nbytes = _ffi.lib().xa_clip_adam_workspace_bytes(0)
    return torch.zeros((nbytes + 7) // 8, dtype=torch.float64, device=device)


def grad_sumsq(grads, workspace, *, stream=None):
    g = _dev(grads, 'float32')
    _call(g, 'xa_grad_sumsq_f32', _ptr(g), g.size, workspace.data_ptr(), workspace.numel() * 8, stream)
    _count()


def clip_adam(param, grad, m, v, step, *, workspace=None, lr=7e-4, beta1=0.9, beta2=0.999, eps=1e-7, clip_norm=None,
              grad_scale=1.0, stream=None):
    """tf.clip_by_global_norm + Keras Adam on one flat fp32 buffer, in place (ppo/agent.py:135-137)."""
    p, g, mm, vv = _dev(param, 'float32'), _dev(grad, 'float32'), _dev(m, 'float32'), _dev(v, 'float32')
    clip = float(clip_norm) if clip_norm else 0.0
    if clip > 0.0:
        if workspace is None:
            raise ValueError('clip_norm needs the workspace that grad_sumsq filled')
        grad_sumsq(grad, workspace, stream=stream)
    _call(p, 'xa_clip_adam_f32', _ptr(p), _ptr(g), _ptr(mm), _ptr(vv), p.size, _tptr(workspace), float(lr), float(beta1),
              float(beta2), float(eps), clip, int(step), float(grad_scale), stream)
    _count()


def grad_from_outputs(d_actor, d_values, grad, *, stream=None):
    """Benchmark stand-in for the network backward: fills the flat gradient from the loss's output gradients."""
    da, dv, g = _dev(d_actor, 'float32'), _dev(d_values, 'float32'), _dev(grad, 'float32')
    _call(g, 'xa_grad_from_outputs_f32', _ptr(da), _ptr(dv), dv.size, da.size // dv.size, _ptr(g), g.size, stream)
    _count()


# ------------------------------------------------------------------------------------------ tensor-core GEMM
def gemm_bf16_tn(a, b, *, bias=None, relu=False, relu_mask=None, out_dtype=torch.float32, out=None, split_k=True, col_group=None,
                 stream=None):
    """C[M,N] = A[M,K] @ B[N,K]^T (+ bias) (ReLU) on tcgen05 tensor cores: bf16 operands, fp32 accumulate.
    The Dense layers of the policy/value network (utils/common.py:239-258), y = x W^T + b, and their backward
    products; shapes with few output tiles and a long K (weight gradients) are split along K over the SMs and
    reduced in a second, deterministic pass.  `relu_mask` [M,N] bf16 zeroes C where mask <= 0.  `col_group` =
    (group, pitch) stores product column j at out[:, (j // group) * pitch + j % group] (`out` required): a [B, h*w*c]
    gradient written straight onto a zero-bordered [B, H, W, c] grid."""
    aa, bb = _dev(a, 'bfloat16'), _dev(b, 'bfloat16')
    (m, k), (n, k2) = aa.shape, bb.shape
    if k != k2:
        raise ValueError(f'inner dimensions differ: A {aa.shape}, B {bb.shape}')
    dev = _device_of(aa)
    c = out if out is not None else torch.empty((m, n), dtype=out_dtype, device=dev)
    bias_a = _dev(bias, 'float32') if bias is not None else None
    mask_a = _dev(relu_mask, 'bfloat16') if relu_mask is not None else None
    ws, ws_bytes = None, 0
    if split_k and mask_a is None and col_group is None:
        ws_bytes = _ffi.lib().xa_gemm_workspace_bytes(m, n, k)
        if ws_bytes:
            ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=dev)
    if col_group is not None and out is None:
        raise ValueError('col_group needs a preallocated `out` (the zero-bordered grid)')
    group, pitch = col_group if col_group is not None else (0, 0)
    _call(aa, 'xa_gemm_bf16_tn_ex', _ptr(aa), _ptr(bb), _tptr(c), _ptr(bias_a), m, n, k, c.stride(0), int(c.dtype == torch.bfloat16),
              int(bool(relu)), _ptr(mask_a), n, int(group), int(pitch), _tptr(ws), ws_bytes, stream)
    _count(2 if ws is not None else 1)
    return c


def to_bf16(src, *, transpose=False, pad_to=8, stream=None):
    """fp32|bf16 [R, C] -> bf16 [R, C] or, transposed, [C, R'] with R' = R rounded up to `pad_to` (zero-filled),
    which is what xa_gemm_bf16_tn needs when R becomes the contraction dimension."""
    a = _dev(src)
    if a.dtype not in ('float32', 'bfloat16') or len(a.shape) != 2:
        raise TypeError(f'to_bf16 needs a 2-D float32/bfloat16 tensor, got {a.dtype} {a.shape}')
    rows, cols = a.shape
    dev = _device_of(a)
    if transpose:
        ld = -(-rows // pad_to) * pad_to
        dst = (torch.zeros if ld != rows else torch.empty)((cols, ld), dtype=torch.bfloat16, device=dev)
    else:
        ld = cols
        dst = torch.empty((rows, cols), dtype=torch.bfloat16, device=dev)
    _call(a, 'xa_to_bf16', _ptr(a), int(a.dtype == 'float32'), _tptr(dst), rows, cols, ld, int(transpose), stream)
    _count()
    return dst


def conv2d_nhwc_bf16(x, w, kh, kw, *, bias=None, relu=False, out_s2d=False, pad=(0, 0), relu_mask=None, out=None, out_hw=None,
                     unpack_s2d=False, zero_border=False, relu_bits_out=None, stream=None):
    """Stride-1 NHWC convolution on tcgen05 (implicit GEMM, no im2col).  x [B,H,W,C] bf16; w [N, kh*kw*C] bf16
    with K ordered (kh, kw, c); zero padding `pad`=(py, px).  Returns y [B,OH,OW,N] bf16, or its 2x2
    space-to-depth form [B,OH/2,OW/2,4N].  `relu_mask` (compact [B,OH,OW,N]) zeroes y where mask <= 0.
    Backward-pass options: `out_hw` computes only that top-left corner of the padded output; a preallocated `out`
    [B,GH,GW,N] larger than the output is a zero-bordered grid written in place; `unpack_s2d` spreads the N = (dy,dx,N/4)
    channels over pixels (2y+dy, 2x+dx) of `out` [B,GH,GW,N/4]; `zero_border` promises that the last pad columns / rows
    of every input image are zero (a gradient on a zero-bordered grid), which lets a padded convolution use the flat
    kernel that fetches every input pixel once.
    ReLU derivatives as bit masks: `relu_bits_out` (int32 [numel(y) / 32], a ReLU layer with a compact or 2x2-packed output)
    receives one bit per output element (> 0); a `relu_mask` of dtype int32 is read as such a bit mask of the compact
    [B,OH,OW,N] tensor instead of the bf16 activation."""
    xx, ww = _dev(x, 'bfloat16'), _dev(w, 'bfloat16')
    B, H, W, C = xx.shape
    N = ww.shape[0]
    if ww.shape[1] != kh * kw * C:
        raise ValueError(f'weights {ww.shape} do not match kh*kw*C = {kh * kw * C}')
    OH, OW = out_hw if out_hw is not None else (H + 2 * pad[0] - kh + 1, W + 2 * pad[1] - kw + 1)
    shape = (B, OH // 2, OW // 2, 4 * N) if out_s2d else ((B, 2 * OH, 2 * OW, N // 4) if unpack_s2d else (B, OH, OW, N))
    y = out if out is not None else torch.empty(shape, dtype=torch.bfloat16, device=_device_of(xx))
    if y.shape[0] != B or y.shape[-1] != shape[-1] or y.shape[1] < shape[1] or y.shape[2] < shape[2] or not y.is_contiguous():
        raise ValueError(f'out {tuple(y.shape)} cannot hold an output of {shape}')
    gh, gw = (y.shape[1], y.shape[2]) if tuple(y.shape) != shape else (0, 0)
    bias_a = _dev(bias, 'float32') if bias is not None else None
    mask_bits = relu_mask is not None and isinstance(relu_mask, torch.Tensor) and relu_mask.dtype == torch.int32
    mask_a = _dev(relu_mask, 'int32' if mask_bits else 'bfloat16') if relu_mask is not None else None
    if mask_a is not None and mask_a.size != B * OH * OW * N // (32 if mask_bits else 1):
        raise ValueError('relu_mask must have the compact output layout [B, OH, OW, N] (one bit per element for an int32 mask)')
    if relu_bits_out is not None and (relu_bits_out.dtype != torch.int32 or relu_bits_out.numel() != y.numel() // 32 or not relu_bits_out.is_contiguous()):
        raise ValueError(f'relu_bits_out must be a contiguous int32 tensor of {y.numel() // 32} words')
    _call(xx, 'xa_conv2d_nhwc_bf16_ex', _ptr(xx), _ptr(ww), _ptr(bias_a), _tptr(y), B, H, W, C, kh, kw, N, int(pad[0]), int(pad[1]),
              int(bool(relu)), 2 if unpack_s2d else int(bool(out_s2d)), _ptr(mask_a), int(OH) if out_hw is not None else 0,
              int(OW) if out_hw is not None else 0, int(gh), int(gw), _tptr(relu_bits_out), int(bool(zero_border)) | (2 if mask_bits else 0), stream)
    _count()
    return y


def conv2d_u8_s2d_bf16(frames, w, kh, kw, *, bias=None, relu=False, out_s2d=False, out=None, x_s2d_out=None, idx=None, time_major=None,
                       relu_bits_out=None, stream=None):
    """The 4x4-strided first layer straight from uint8 frames [B,H,W,4]: /255, space-to-depth and the kh x kw stride-1
    convolution over the [B,H/4,W/4,64] grid in one kernel (bit-identical to space_to_depth_u8_bf16 + conv2d_nhwc_bf16).
    `x_s2d_out` [B,H/4,W/4,64] bf16 also receives the scaled space-to-depth tensor (for the weight gradient).
    `idx` (int32 [B]): the layer runs on frames[idx] without gathering them (`frames` is then the whole store; with
    `time_major=(T, E)` the ids are env-major sample ids of a time-major rollout, as for `gather_rows`)."""
    f, ww = _dev(frames, 'uint8'), _dev(w, 'bfloat16')
    n_frames, H, W, C = f.shape
    ids = _dev(idx, 'int32') if idx is not None else None
    B = ids.shape[0] if ids is not None else n_frames
    N = ww.shape[0]
    if C != 4 or ww.shape[1] != kh * kw * 64:
        raise ValueError(f'frames {f.shape} / weights {ww.shape}: built for 4-channel frames and K = kh*kw*64')
    OH, OW = H // 4 - kh + 1, W // 4 - kw + 1
    shape = (B, OH // 2, OW // 2, 4 * N) if out_s2d else (B, OH, OW, N)
    y = out if out is not None else torch.empty(shape, dtype=torch.bfloat16, device=_device_of(f))
    bias_a = _dev(bias, 'float32') if bias is not None else None
    if x_s2d_out is not None and (x_s2d_out.dtype != torch.bfloat16 or tuple(x_s2d_out.shape) != (B, H // 4, W // 4, 64) or not x_s2d_out.is_contiguous()):
        raise ValueError(f'x_s2d_out must be a contiguous bf16 [{B}, {H // 4}, {W // 4}, 64] tensor')
    if relu_bits_out is not None and (relu_bits_out.dtype != torch.int32 or relu_bits_out.numel() != y.numel() // 32 or not relu_bits_out.is_contiguous()):
        raise ValueError(f'relu_bits_out must be a contiguous int32 tensor of {y.numel() // 32} words')
    if ids is not None or relu_bits_out is not None:
        T, E = _layout(time_major)
        _call(f, 'xa_conv2d_u8_s2d_bf16_ex', _ptr(f), n_frames, _ptr(ids), T, E, _ptr(ww), _ptr(bias_a), _tptr(y), _tptr(x_s2d_out),
              _tptr(relu_bits_out), B, H, W, kh, kw, N, int(bool(relu)), int(bool(out_s2d)), stream)
    else:
        _call(f, 'xa_conv2d_u8_s2d_bf16', _ptr(f), _ptr(ww), _ptr(bias_a), _tptr(y), _tptr(x_s2d_out), B, H, W, kh, kw, N, int(bool(relu)),
              int(bool(out_s2d)), stream)
    _count()
    return y


def space_to_depth_u8_bf16(frames, block, *, scale_255=True, out=None, stream=None):
    """uint8 [B,H,W,C] -> bf16 [B,H/s,W/s,s*s*C] (channel order dy, dx, c), optionally divided by 255."""
    f = _dev(frames, 'uint8')
    B, H, W, C = f.shape
    y = out if out is not None else torch.empty((B, H // block, W // block, block * block * C), dtype=torch.bfloat16,
                                                 device=_device_of(f))
    _call(f, 'xa_space_to_depth_u8_bf16', _ptr(f), _tptr(y), B, H, W, C, block, int(bool(scale_255)), stream)
    _count()
    return y


def gather_s2d_u8_bf16(frames, idx, block=4, *, time_major=None, scale_255=True, out=None, stream=None):
    """Minibatch gather + /255 + space-to-depth in one pass: frames [T,E,H,W,C] (time_major=(T,E)) or [rows,H,W,C] uint8,
    idx env-major sample ids -> bf16 [n, H/s, W/s, s*s*C], the first-layer input of the tensor-core network."""
    f, i = _dev(frames, 'uint8'), _dev(idx, 'int32')
    T, E = _layout(time_major)
    H, W, C = f.shape[-3:]
    n_rows = f.shape[0] * (f.shape[1] if T else 1)
    n = i.size
    y = out if out is not None else torch.empty((n, H // block, W // block, block * block * C), dtype=torch.bfloat16,
                                                 device=_device_of(f))
    _call(f, 'xa_gather_s2d_u8_bf16', _ptr(f), _ptr(i), _tptr(y), n, n_rows, T, E, H, W, C, block, int(bool(scale_255)),
              stream)
    _count()
    return y


_wgrad_nhwc_ws = {}


def conv_wgrad_nhwc_bf16(x, dy_grid, kh, kw, *, stream=None):
    """dW [N, kh*kw*C] and db [N] (fp32) of a stride-1 convolution from the natural NHWC tensors: x [B,H,W,C] bf16 and
    dy_grid [B,H,W,N] bf16 = dY on the input grid (zero outside the [OH, OW] corner).  No transposed copies."""
    xx, dd = _dev(x, 'bfloat16'), _dev(dy_grid, 'bfloat16')
    B, H, W, C = xx.shape
    N = dd.shape[-1]
    if tuple(dd.shape[:3]) != (B, H, W):
        raise ValueError(f'dy_grid {tuple(dd.shape)} is not on the grid of x {tuple(xx.shape)}')
    dev = _device_of(xx)
    key = (dev, N, C, kh, kw)
    ws = _wgrad_nhwc_ws.get(key)
    if ws is None:
        nbytes = _ffi.lib().xa_conv_wgrad_nhwc_workspace_bytes(N, C, kh, kw)
        ws = _wgrad_nhwc_ws[key] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dw = torch.empty((N, kh * kw * C), dtype=torch.float32, device=dev)
    db = torch.empty((N,), dtype=torch.float32, device=dev)
    _call(xx, 'xa_conv_wgrad_nhwc_bf16', _ptr(xx), _ptr(dd), _tptr(dw), _tptr(db), N, C, kh, kw, W, B * H * W, _tptr(ws),
              ws.numel(), stream)
    _count(2)
    return dw, db


def gemm_bf16_atb(a, b, *, split_k=True, stream=None):
    """C[M,N] fp32 = A^T B for row-major A [K,M] and B [K,N] bf16 -- the Dense weight gradient dW = dY^T X from the
    tensors as the forward / data-gradient kernels leave them (no transposed copies)."""
    aa, bb = _dev(a, 'bfloat16'), _dev(b, 'bfloat16')
    (k, m), (k2, n) = aa.shape, bb.shape
    if k != k2:
        raise ValueError(f'contraction lengths differ: A {aa.shape}, B {bb.shape}')
    dev = _device_of(aa)
    c = torch.empty((m, n), dtype=torch.float32, device=dev)
    ws, ws_bytes = None, 0
    if split_k:
        ws_bytes = _ffi.lib().xa_gemm_atb_workspace_bytes(m, n, k)
        if ws_bytes:
            ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=dev)
    _call(aa, 'xa_gemm_bf16_atb', _ptr(aa), _ptr(bb), _tptr(c), m, n, k, n, _tptr(ws), ws_bytes, stream)
    _count(2 if ws is not None else 1)
    return c


def gather_cast_f32(src, index_map, out, *, stream=None):
    """out[i] = src[index_map[i]] (0 where the index is negative), cast to out's dtype (bf16 or fp32)."""
    s_, m_ = _dev(src, 'float32'), _dev(index_map, 'int32')
    if out.dtype not in (torch.bfloat16, torch.float32) or out.numel() != m_.size:
        raise ValueError('out must be a bf16/fp32 tensor with one element per map entry')
    _call(s_, 'xa_gather_cast_f32', _ptr(s_), _ptr(m_), _tptr(out), out.numel(), int(out.dtype == torch.bfloat16), stream)
    _count()
    return out
