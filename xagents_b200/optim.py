"""tf.clip_by_global_norm + Keras Adam over one flat fp32 buffer with the launch arguments resolved once.

`ops.clip_adam` converts five tensors through DLPack on every call (~40 us of host time for the two launches); an update
happens after every minibatch, so the agents' models and bench.py hold a `FlatAdam` instead: fixed device pointers,
two ctypes calls per update (xa_grad_sumsq_f32, xa_clip_adam_f32 -- csrc/optim.cu; reference: xagents/ppo/agent.py:135-137,
Adam built at xagents/utils/common.py:476).
"""
import ctypes

import torch

from . import _ffi, ops


class FlatAdam:
    def __init__(self, param, grad, lr=7e-4, beta1=0.9, beta2=0.999, eps=1e-7):
        assert param.is_cuda and param.dtype == torch.float32 and param.is_contiguous() and param.numel() % 4 == 0
        assert grad.shape == param.shape and grad.dtype == torch.float32 and grad.device == param.device
        self.param, self.grad = param, grad
        self.m, self.v = torch.zeros_like(param), torch.zeros_like(param)
        self.workspace = ops.optim_workspace(param.device)
        self.lr, self.beta1, self.beta2, self.eps = float(lr), float(beta1), float(beta2), float(eps)
        self.device = param.device
        self.t = 0
        lib = _ffi.lib()
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        self._sumsq, self._adam = lib.xa_grad_sumsq_f32, lib.xa_clip_adam_f32
        self._sumsq_args = (p(grad), param.numel(), p(self.workspace), self.workspace.numel() * 8)
        self._adam_args = (p(param), p(grad), p(self.m), p(self.v), param.numel(), p(self.workspace))

    def step(self, clip_norm=None, grad_scale=1.0, stream=None):
        """One update from `self.grad` (times `grad_scale`: 1/world after a sum all-reduce), in place, asynchronous."""
        self.t += 1
        s = ctypes.c_void_p((stream if stream is not None else torch.cuda.current_stream(self.device)).cuda_stream)
        clip = float(clip_norm) if clip_norm else 0.0
        if self.device.index != torch.cuda.current_device():
            with torch.cuda.device(self.device):
                return self._launch(clip, float(grad_scale), s)
        return self._launch(clip, float(grad_scale), s)

    def _launch(self, clip, scale, s):
        if clip > 0.0:
            _ffi.check('xa_grad_sumsq_f32', self._sumsq(*self._sumsq_args, s))
            ops._count()
        _ffi.check('xa_clip_adam_f32', self._adam(*self._adam_args, self.lr, self.beta1, self.beta2, self.eps, clip, self.t, scale, s))
        ops._count()
