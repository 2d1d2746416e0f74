"""ctypes binding of include/xagents_b200.h (the only way the Python layer reaches the kernels).

There is deliberately no fallback: if the shared library is missing, or a call returns non-zero, an
exception is raised.  Build it with `python -m xagents_b200._build` (or `__graft_entry__.build()`).
"""
import ctypes
import os

from ._build import LIB_PATH

c_f32p = ctypes.c_void_p      # device pointers travel as integers
c_stream = ctypes.c_void_p

XA_SCAN_AUTO, XA_SCAN_SEQUENTIAL, XA_SCAN_CHUNKED = 0, 1, 2
XA_GATHER_AUTO, XA_GATHER_BULK, XA_GATHER_VECTOR = 0, 1, 2
XA_ACTOR_LOGITS, XA_ACTOR_PROBS, XA_ACTOR_NORMAL = 0, 1, 2
XA_MAX_FIELDS = 8
XA_MOMENT_STRIDE = 4


class XAError(RuntimeError):
    def __init__(self, fn, code, message):
        super().__init__(f'{fn} failed with code {code}: {message}')
        self.fn, self.code, self.message = fn, code, message


class LossArgs(ctypes.Structure):
    """struct xa_loss_args"""
    _fields_ = [
        ('actor_out', ctypes.c_void_p), ('values', ctypes.c_void_p), ('actions', ctypes.c_void_p),
        ('old_log_probs', ctypes.c_void_p), ('old_values', ctypes.c_void_p), ('returns', ctypes.c_void_p),
        ('idx', ctypes.c_void_p), ('n_steps', ctypes.c_int32), ('n_envs', ctypes.c_int32),
        ('advantages', ctypes.c_void_p), ('moments', ctypes.c_void_p), ('n_moment_parts', ctypes.c_int32),
        ('moment_part_stride', ctypes.c_int64), ('n', ctypes.c_int64), ('n_actions', ctypes.c_int32),
        ('actor_kind', ctypes.c_int32), ('clip', ctypes.c_float), ('ent_coef', ctypes.c_float),
        ('vf_coef', ctypes.c_float), ('adv_eps', ctypes.c_float), ('out_scalars', ctypes.c_void_p),
        ('d_actor', ctypes.c_void_p), ('d_values', ctypes.c_void_p), ('advantages_out', ctypes.c_void_p),
        ('workspace', ctypes.c_void_p), ('workspace_bytes', ctypes.c_int64),
    ]


XA_MAX_PEERS = 8


class PeerAdamArgs(ctypes.Structure):
    """struct xa_peer_adam_args"""
    _fields_ = [('grad', ctypes.c_void_p * XA_MAX_PEERS), ('param', ctypes.c_void_p * XA_MAX_PEERS),
                ('sumsq', ctypes.c_void_p * XA_MAX_PEERS), ('flags', ctypes.c_void_p * XA_MAX_PEERS),
                ('m', ctypes.c_void_p), ('v', ctypes.c_void_p), ('workspace', ctypes.c_void_p), ('shard_len', ctypes.c_int64),
                ('rank', ctypes.c_int32), ('world', ctypes.c_int32), ('epoch', ctypes.c_uint32), ('lr_t', ctypes.c_float),
                ('beta1', ctypes.c_float), ('beta2', ctypes.c_float), ('eps', ctypes.c_float), ('clip', ctypes.c_float),
                ('inv_world', ctypes.c_float)]


class PeerGatherArgs(ctypes.Structure):
    """struct xa_peer_gather_args"""
    _fields_ = [('recv', ctypes.c_void_p * XA_MAX_PEERS), ('flags', ctypes.c_void_p * XA_MAX_PEERS), ('status', ctypes.c_void_p),
                ('n_per_rank', ctypes.c_int64), ('rank', ctypes.c_int32), ('world', ctypes.c_int32), ('epoch', ctypes.c_uint32)]


XA_MAX_GRAD_SEGMENTS = 16


class GradSegment(ctypes.Structure):
    """struct xa_grad_segment_t"""
    _fields_ = [('dest_begin', ctypes.c_int64), ('split_stride', ctypes.c_int64), ('splits', ctypes.c_int32), ('wide', ctypes.c_int32)]


class NatureCnn(ctypes.Structure):
    """struct xa_nature_cnn_t"""
    _fields_ = ([('batch', ctypes.c_int32), ('n_actions', ctypes.c_int32)] +
                [(name, ctypes.c_void_p) for name in ('w1', 'w2', 'w3', 'w2_flip', 'w3_flip', 'wf', 'wf_t', 'wh', 'b1', 'b2', 'b3', 'bf', 'bh',
                                                      'x1', 'x2', 'x3', 'y3', 'h', 'actor', 'critic', 'dh', 'g3', 'g2', 'g1', 'gemm_ws')] +
                [('gemm_ws_bytes', ctypes.c_int64), ('scratch', ctypes.c_void_p), ('scratch_floats', ctypes.c_int64)] +
                [(name, ctypes.c_int64) for name in ('off_c1', 'off_c2', 'off_c3', 'off_fc', 'off_heads')] +
                [('grad_map', ctypes.c_void_p), ('grad_dest', ctypes.c_void_p), ('segments', GradSegment * XA_MAX_GRAD_SEGMENTS), ('n_segments', ctypes.c_int32),
                 ('reserved', ctypes.c_int32), ('n_grad', ctypes.c_int64), ('relu_bits2', ctypes.c_void_p), ('relu_bits3', ctypes.c_void_p),
                 ('relu_bitsf', ctypes.c_void_p)])


# name -> (restype, argtypes); every symbol include/xagents_b200.h declares
PROTOTYPES = {
    'xa_version': (ctypes.c_int, []),
    'xa_last_error': (ctypes.c_char_p, []),
    'xa_device_info': (ctypes.c_int, [ctypes.c_int] + [ctypes.POINTER(ctypes.c_int)] * 3),
    'xa_set_chained_launches': (ctypes.c_int, [ctypes.c_int]),
    'xa_gae_f32': (ctypes.c_int, [c_f32p] * 6 + [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                                  ctypes.c_int, c_stream]),
    'xa_nstep_returns_f32': (ctypes.c_int, [c_f32p] * 4 + [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                                            c_stream]),
    'xa_retrace_f32': (ctypes.c_int, [c_f32p] * 7 + [ctypes.c_int, ctypes.c_int, ctypes.c_double, c_stream]),
    'xa_gather_rows': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                      ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_stream]),
    'xa_gather_fields_f32': (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p), ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, c_stream]),
    'xa_gather_minibatch': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                           ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p), ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           c_stream]),
    'xa_gather_progress_units': (ctypes.c_int, [ctypes.c_int64]),
    'xa_gather_rows_progress': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                               ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, c_stream]),
    'xa_wait_progress_u32': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, c_stream]),
    'xa_stream_wait_geq_u32': (ctypes.c_int, [c_stream, ctypes.c_void_p, ctypes.c_uint32]),
    'xa_gather_rows_u8_scaled_f32': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                                    ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, c_stream]),
    'xa_adv_moments_f32': (ctypes.c_int, [c_f32p, c_f32p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_void_p, c_stream]),
    'xa_normalize_adv_f32': (ctypes.c_int, [c_f32p, c_f32p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double, c_f32p, c_stream]),
    'xa_loss_workspace_bytes': (ctypes.c_int64, [ctypes.c_int64]),
    'xa_ppo_loss_f32': (ctypes.c_int, [ctypes.POINTER(LossArgs), c_stream]),
    'xa_a2c_loss_f32': (ctypes.c_int, [ctypes.POINTER(LossArgs), c_stream]),
    'xa_policy_step_f32': (ctypes.c_int, [c_f32p, ctypes.c_int, c_f32p, ctypes.c_uint64, ctypes.c_uint64, c_f32p, c_f32p, c_f32p,
                                          ctypes.c_int64, ctypes.c_int, c_stream]),
    'xa_policy_step_counter_f32': (ctypes.c_int, [c_f32p, ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, c_f32p, c_f32p,
                                                  c_f32p, ctypes.c_int64, ctypes.c_int, c_stream]),
    'xa_bump_u64': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, c_stream]),
    'xa_synth_env_step_u8': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, c_f32p, c_f32p, c_f32p, c_f32p,
                                            ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, c_stream]),
    'xa_gemm_bf16_tn': (ctypes.c_int, [ctypes.c_void_p] * 3 + [c_f32p] + [ctypes.c_int64] * 4 + [ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                                                                ctypes.c_void_p, ctypes.c_int64, c_stream]),
    'xa_gemm_bf16_tn_ex': (ctypes.c_int, [ctypes.c_void_p] * 3 + [c_f32p] + [ctypes.c_int64] * 4 + [ctypes.c_int, ctypes.c_int, ctypes.c_void_p] +
                           [ctypes.c_int64] * 3 + [ctypes.c_void_p, ctypes.c_int64, c_stream]),
    'xa_conv2d_nhwc_bf16_ex': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_f32p, ctypes.c_void_p] + [ctypes.c_int] * 11 +
                               [ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_int, c_stream]),
    'xa_conv2d_u8_s2d_bf16_ex': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, c_f32p,
                                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 8 + [c_stream]),
    'xa_gather_cast_f32': (ctypes.c_int, [c_f32p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, c_stream]),
    'xa_gemm_atb_workspace_bytes': (ctypes.c_int64, [ctypes.c_int64] * 3),
    'xa_gemm_bf16_atb': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_f32p] + [ctypes.c_int64] * 4 + [ctypes.c_void_p, ctypes.c_int64, c_stream]),
    'xa_gemm_workspace_bytes': (ctypes.c_int64, [ctypes.c_int64] * 3),
    'xa_gemm_bf16_tn_maskbits': (ctypes.c_int, [ctypes.c_void_p] * 3 + [ctypes.c_int64] * 4 + [ctypes.c_void_p] + [ctypes.c_int64] * 3 + [c_stream]),
    'xa_gemm_bf16_tn_partial': (ctypes.c_int, [ctypes.c_void_p] * 2 + [ctypes.c_int64] * 3 + [ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int), c_stream]),
    'xa_conv2d_nhwc_bf16': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_f32p, ctypes.c_void_p] + [ctypes.c_int] * 11 +
                            [ctypes.c_void_p, c_stream]),
    'xa_conv2d_u8_s2d_bf16': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_f32p, ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 8 + [c_stream]),
    'xa_space_to_depth_u8_bf16': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 6 + [c_stream]),
    'xa_conv_wgrad_nhwc_workspace_bytes': (ctypes.c_int64, [ctypes.c_int] * 4),
    'xa_conv_wgrad_nhwc_bf16': (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_int] * 5 + [ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, c_stream]),
    'xa_conv_wgrad_nhwc_plan': (ctypes.c_int, [ctypes.c_int] * 4 + [ctypes.c_int64, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    'xa_conv_wgrad_nhwc_bf16_partial': (ctypes.c_int, [ctypes.c_void_p] * 2 + [ctypes.c_int] * 5 + [ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, c_stream]),
    'xa_gemm_atb_plan': (ctypes.c_int, [ctypes.c_int64] * 3 + [ctypes.POINTER(ctypes.c_int)]),
    'xa_gemm_bf16_atb_partial': (ctypes.c_int, [ctypes.c_void_p] * 2 + [ctypes.c_int64] * 3 + [ctypes.c_void_p, ctypes.c_int64, c_stream]),
    'xa_heads_backward_blocks': (ctypes.c_int, [ctypes.c_int]),
    'xa_heads_forward_bf16': (ctypes.c_int, [ctypes.c_void_p] * 5 + [ctypes.c_int] * 3 + [c_stream]),
    'xa_heads_forward_partial_bf16': (ctypes.c_int, [c_f32p, ctypes.c_int, c_f32p, ctypes.c_void_p, ctypes.c_void_p, c_f32p, c_f32p, c_f32p] + [ctypes.c_int] * 3 + [c_stream]),
    'xa_heads_backward_bf16': (ctypes.c_int, [ctypes.c_void_p] * 6 + [ctypes.c_int64] + [ctypes.c_int] * 3 + [c_stream]),
    'xa_grad_finalize_f32': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(GradSegment), ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_int64, c_stream]),
    'xa_nature_cnn_forward': (ctypes.c_int, [ctypes.POINTER(NatureCnn), ctypes.c_void_p, ctypes.c_int, c_stream]),
    'xa_nature_cnn_forward_indexed': (ctypes.c_int, [ctypes.POINTER(NatureCnn), ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                     c_stream]),
    'xa_conv2d_u8_s2d_bf16_indexed': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, c_f32p,
                                                     ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 8 + [c_stream]),
    'xa_nature_cnn_backward': (ctypes.c_int, [ctypes.POINTER(NatureCnn)] + [ctypes.c_void_p] * 4 + [c_stream]),
    'xa_gather_s2d_u8_bf16': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64] +
                              [ctypes.c_int] * 7 + [c_stream]),
    'xa_to_bf16': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                  ctypes.c_int, c_stream]),
    'xa_clip_adam_workspace_bytes': (ctypes.c_int64, [ctypes.c_int64]),
    'xa_grad_sumsq_f32': (ctypes.c_int, [c_f32p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, c_stream]),
    'xa_ipc_alloc': (ctypes.c_int, [ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p]),
    'xa_ipc_open': (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
    'xa_ipc_close': (ctypes.c_int, [ctypes.c_void_p]),
    'xa_ipc_free': (ctypes.c_int, [ctypes.c_void_p]),
    'xa_peer_allgather_f64': (ctypes.c_int, [ctypes.POINTER(PeerGatherArgs), ctypes.c_void_p, ctypes.c_void_p, c_stream]),
    'xa_peer_adam_workspace_bytes': (ctypes.c_int64, []),
    'xa_peer_adam_flag_bytes': (ctypes.c_int64, []),
    'xa_peer_allreduce_adam_f32': (ctypes.c_int, [ctypes.POINTER(PeerAdamArgs)] + [ctypes.c_double] * 5 + [ctypes.c_int64, c_stream]),
    'xa_grad_from_outputs_f32': (ctypes.c_int, [c_f32p, c_f32p, ctypes.c_int64, ctypes.c_int, c_f32p, ctypes.c_int64, c_stream]),
    'xa_clip_adam_f32': (ctypes.c_int, [c_f32p] * 4 + [ctypes.c_int64, ctypes.c_void_p, ctypes.c_double, ctypes.c_double,
                                                       ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int64,
                                                       ctypes.c_double, c_stream]),
}

_lib = None


def library_path():
    return os.environ.get('XAGENTS_B200_LIB', LIB_PATH)


def lib():
    """The loaded shared library (loads on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise ImportError(f'{path} is missing: run `python -m xagents_b200._build` (nvcc, sm_100a). '
                              'xagents_b200 has no CPU or framework fallback.')
        handle = ctypes.CDLL(path)
        for name, (restype, argtypes) in PROTOTYPES.items():
            fn = getattr(handle, name)          # AttributeError here = header and library out of sync
            fn.restype, fn.argtypes = restype, argtypes
        _lib = handle
    return _lib


def check(fn_name, code):
    if code != 0:
        raise XAError(fn_name, code, lib().xa_last_error().decode('utf-8', 'replace'))


def call(fn_name, *args):
    """Call an int-returning entry point and raise XAError on a non-zero code."""
    check(fn_name, getattr(lib(), fn_name)(*args))
