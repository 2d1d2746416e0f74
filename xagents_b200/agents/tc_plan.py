"""The Nature CNN's forward / backward as TWO native calls per minibatch (csrc/nature_net.cu) over buffers with fixed addresses.

`NaturePlan` owns, for one batch size, every activation and gradient tensor of the tensor-core network, the fp32 scratch its
split partial sums land in, and the index map + segment table of `xa_grad_finalize_f32`, the last launch of the backward
pass, which adds the partials and writes every parameter gradient at its place in the model's flat gradient buffer (torch
layouts) -- so one train-step minibatch is `forward()` (6 launches: the first layer reads the uint8 frames directly), the loss kernel, `backward()` (9 launches) and the fused
clip + Adam (xagents/ppo/agent.py:112-137: model call, tape.gradient, clip_by_global_norm, apply_gradients).

The autograd form of the same pipeline (agents/tc_cnn.py `_NatureCnnFn`) stays as the checker: tests compare the two.
"""
import ctypes
import os

import torch

from .. import _ffi, ops
from .tc_operands import derive

HIDDEN, HEAD_ROWS = 512, 8


class NaturePlan:
    def __init__(self, pack, named, flat, batch, n_grad=None, backward=True):
        """pack: OperandPack; named: its name -> parameter dict; flat: the flat fp32 buffer the parameters are views of (their
        offsets in it are the offsets of their gradients in the flat gradient buffer); n_grad: length of that buffer."""
        lib = _ffi.lib()
        dev = pack.w1.device
        self.device, self.batch, self.pack = dev, int(batch), pack
        B, A = self.batch, pack.n_actions
        bf16 = lambda *shape: torch.empty(shape, dtype=torch.bfloat16, device=dev)
        self.x1 = None                                            # allocated on first use with uint8 frames
        self.x2, self.x3, self.y3, self.h = bf16(B, 10, 10, 128), bf16(B, 9, 9, 64), bf16(B, 7, 7, 64), bf16(B, HIDDEN)
        self.actor = torch.empty((B, A), dtype=torch.float32, device=dev)
        self.critic = torch.empty((B,), dtype=torch.float32, device=dev)
        ws_bytes = lib.xa_gemm_workspace_bytes(B, HIDDEN, 3136)
        self.gemm_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
        net = self.net = _ffi.NatureCnn()
        net.batch, net.n_actions = B, A
        for name in ('w1', 'w2', 'w3', 'w2_flip', 'w3_flip', 'wf', 'wf_t', 'wh', 'b1', 'b2', 'b3', 'bh'):
            setattr(net, name, getattr(pack, name).data_ptr())
        net.bf = pack.bf_.data_ptr()
        for name in ('x2', 'x3', 'y3', 'h', 'actor', 'critic'):
            setattr(net, name, getattr(self, name).data_ptr())
        if self.gemm_ws is not None:
            net.gemm_ws, net.gemm_ws_bytes = self.gemm_ws.data_ptr(), ws_bytes
        self._x1_in = None                                        # the first-layer input of the last forward (for the backward)
        self.has_backward = bool(backward)
        if backward:
            self._build_backward(lib, named, flat, n_grad)
        self._fwd, self._bwd, self._fwd_idx = lib.xa_nature_cnn_forward, lib.xa_nature_cnn_backward, lib.xa_nature_cnn_forward_indexed

    # ---- backward buffers, scratch layout, gradient map -----------------------------------------------------------
    def _build_backward(self, lib, named, flat, n_grad):
        dev, B, A, net = self.device, self.batch, self.pack.n_actions, self.net
        zeros = lambda *shape: torch.zeros(shape, dtype=torch.bfloat16, device=dev)   # only the valid corners are ever written
        self.dh = torch.empty((B, HIDDEN), dtype=torch.bfloat16, device=dev)
        self.g3, self.g2, self.g1 = zeros(B, 9, 9, 64), zeros(B, 10, 10, 64), zeros(B, 21, 21, 32)
        for name in ('dh', 'g3', 'g2', 'g1'):
            setattr(net, name, getattr(self, name).data_ptr())
        # ReLU derivatives as bit masks: written by the forward epilogues of conv1 / conv2 (one bit per element of x2 / x3), read by
        # the data gradients of conv2 / conv3 instead of the bf16 activations (295 MB -> 18 MB per 8192-frame minibatch)
        self.bits2 = torch.empty(B * 10 * 10 * 128 // 32, dtype=torch.int32, device=dev)
        self.bits3 = torch.empty(B * 9 * 9 * 64 // 32, dtype=torch.int32, device=dev)
        if os.environ.get('XA_NO_RELU_BITS') != '1':               # tuning switch: the data gradients read the bf16 activations instead
            net.relu_bits2, net.relu_bits3 = self.bits2.data_ptr(), self.bits3.data_ptr()
            self.bitsf = torch.empty(B * 3136 // 32, dtype=torch.int32, device=dev)      # y3's signs, for the FC layer's data gradient
            net.relu_bitsf = self.bitsf.data_ptr()
        splits, ld = ctypes.c_int(), ctypes.c_int()
        conv = {}
        off = 0
        for key, (n_out, ch, kh, kw, q) in dict(c1=(32, 64, 2, 2, B * 441), c2=(64, 128, 2, 2, B * 100), c3=(64, 64, 3, 3, B * 81)).items():
            _ffi.check('xa_conv_wgrad_nhwc_plan', lib.xa_conv_wgrad_nhwc_plan(n_out, ch, kh, kw, q, ctypes.byref(splits), ctypes.byref(ld)))
            room = lib.xa_conv_wgrad_nhwc_workspace_bytes(n_out, ch, kh, kw) // 4
            conv[key] = (off, splits.value, ld.value, n_out, kh * kw * ch)
            off += -(-room // 64) * 64
        _ffi.check('xa_gemm_atb_plan', lib.xa_gemm_atb_plan(HIDDEN, 3136, B, ctypes.byref(splits)))
        fc_off, fc_splits = off, splits.value
        off += fc_splits * HIDDEN * 3136
        heads_off, heads_blocks = off, lib.xa_heads_backward_blocks(B)
        off += heads_blocks * (HEAD_ROWS + 2) * HIDDEN
        self.scratch = torch.empty(off, dtype=torch.float32, device=dev)
        net.scratch, net.scratch_floats = self.scratch.data_ptr(), off
        net.off_c1, net.off_c2, net.off_c3, net.off_fc, net.off_heads = conv['c1'][0], conv['c2'][0], conv['c3'][0], fc_off, heads_off

        # where parameter element j's gradient sits in `scratch` (split 0): the operand layouts of tc_operands.derive applied to
        # index tensors give, for every operand element, the flat index of the parameter it is a copy of
        offsets = {name: (q.data_ptr() - flat.data_ptr()) // 4 for name, q in named.items()}
        n_params = max(offsets[name] + q.numel() for name, q in named.items())
        n_grad = int(n_grad) if n_grad is not None else n_params
        assert sorted(offsets.values())[0] == 0 and n_grad >= n_params, 'the parameters must tile the flat buffer from its start'
        ids = {name: torch.arange(offsets[name], offsets[name] + q.numel(), dtype=torch.float64, device=dev).view_as(q) for name, q in named.items()}
        lay16, lay32 = derive(ids, -1.0)
        gmap = torch.full((n_grad,), -1, dtype=torch.int64, device=dev)

        def put(layout_ids, src):
            layout_ids, src = layout_ids.reshape(-1).to(torch.int64), src.reshape(-1).to(torch.int64)
            valid = layout_ids >= 0
            gmap[layout_ids[valid]] = src[valid]

        ar = lambda n: torch.arange(n, dtype=torch.int64, device=dev)
        for key, wname, bname in (('c1', 'w1', 'b1'), ('c2', 'w2', 'b2'), ('c3', 'w3', 'b3')):
            base, _, ld_p, n_out, k = conv[key]
            put(lay16[wname], base + ar(n_out)[:, None] * ld_p + ar(k)[None, :])
            put(lay32[bname], base + ar(n_out) * ld_p + k)
        put(lay16['wf'], fc_off + ar(HIDDEN * 3136))
        put(lay32['bf_'], heads_off + HEAD_ROWS * HIDDEN + ar(HIDDEN))
        rows = lay16['wh'].shape[0]
        assert rows == HEAD_ROWS, f'{self.pack.n_actions} actions need {rows} stacked head rows; the heads kernel holds {HEAD_ROWS}'
        put(lay16['wh'], heads_off + ar(rows)[:, None] * HIDDEN + ar(HIDDEN)[None, :])
        put(lay32['bh'], heads_off + (HEAD_ROWS + 1) * HIDDEN + ar(rows))
        assert bool((gmap[:n_params] >= 0).all()), 'a parameter has no gradient source'
        assert int(gmap.max()) < 2 ** 31
        # enumerate every parameter's elements in SOURCE order (coalesced reads of the partial sums; the one result per
        # element is the scattered store): (grad_map[j], grad_dest[j]) = (source offset, index in flat_grad)
        order = sorted(named, key=lambda name: offsets[name])
        dest = torch.arange(n_grad, dtype=torch.int64, device=dev)
        for i, name in enumerate(order):
            lo = offsets[name]
            hi = offsets[order[i + 1]] if i + 1 < len(order) else n_grad
            assert hi >= lo + named[name].numel(), 'parameters overlap in the flat buffer'
            key = gmap[lo:hi].clone()
            key[key < 0] = 2 ** 40                                   # padding (no source) last
            perm = torch.argsort(key, stable=True)
            dest[lo:hi] = lo + perm
        self.grad_map = gmap[dest].to(torch.int32)
        self.grad_dest = dest.to(torch.int32)
        net.grad_map, net.grad_dest, net.n_grad = self.grad_map.data_ptr(), self.grad_dest.data_ptr(), n_grad

        seg_of = dict(c1w=('c1',), c1b=('c1',), c2w=('c2',), c2b=('c2',), c3w=('c3',), c3b=('c3',))
        assert len(order) <= _ffi.XA_MAX_GRAD_SEGMENTS
        for i, name in enumerate(order):
            seg = net.segments[i]
            seg.dest_begin = offsets[name]
            if name in seg_of:
                _, n_splits, ld_p, n_out, _ = conv[seg_of[name][0]]
                seg.splits, seg.split_stride = n_splits, n_out * ld_p
            elif name == 'fcw':
                seg.splits, seg.split_stride = fc_splits, HIDDEN * 3136
            else:                                                  # fcb, aw, ab, cw, cb: per-CTA blocks of the heads kernel
                seg.splits, seg.split_stride = heads_blocks, (HEAD_ROWS + 2) * HIDDEN
            seg.wide = 0 if seg.splits < 8 else (8 if seg.splits >= 64 and name not in seg_of and name != 'fcw' else 4)
        net.n_segments = len(order)

    # ---- the two calls ------------------------------------------------------------------------------------------
    def _stream(self, stream):
        return ctypes.c_void_p((stream if stream is not None else torch.cuda.current_stream(self.device)).cuda_stream)

    def forward(self, frames, stream=None, out=None, idx=None, time_major=None):
        """uint8 [B,84,84,4] frames, or the bf16 [B,21,21,64] output of ops.gather_s2d_u8_bf16 -> (actor [B,A], critic [B]) fp32:
        the plan's own buffers, valid until the next forward(), or the caller's `out` = (actor, critic) tensors.
        `idx` (int32 [B], device): the batch is frames[idx] of a larger uint8 frame store, read through the permutation by the
        first layer (no gathered copy); `time_major=(T, E)`: ids are env-major sample ids of a time-major rollout."""
        assert frames.is_cuda and frames.is_contiguous() and (idx is not None or frames.shape[0] == self.batch), (tuple(frames.shape), self.batch)
        actor, critic = out if out is not None else (self.actor, self.critic)
        if out is not None:
            assert actor.dtype == torch.float32 and actor.is_contiguous() and actor.numel() == self.actor.numel() and actor.device == self.device
            assert critic.dtype == torch.float32 and critic.is_contiguous() and critic.numel() == self.batch and critic.device == self.device
        self.net.actor, self.net.critic = actor.data_ptr(), critic.data_ptr()
        s2d = frames.dtype == torch.bfloat16
        if s2d:
            assert tuple(frames.shape[1:]) == (21, 21, 64)
            self._x1_in = frames
        else:
            assert frames.dtype == torch.uint8 and tuple(frames.shape[1:]) == (84, 84, 4)
            if self.x1 is None and self.has_backward:           # the first layer reads the frames; only the backward needs x1
                self.x1 = torch.empty((self.batch, 21, 21, 64), dtype=torch.bfloat16, device=self.device)
                self.net.x1 = self.x1.data_ptr()
            self._x1_in = self.x1
        if idx is not None:
            assert not s2d and idx.dtype == torch.int32 and idx.is_cuda and idx.is_contiguous() and idx.numel() == self.batch
            T, E = (int(time_major[0]), int(time_major[1])) if time_major is not None else (0, 0)
            assert T <= 0 or T * E == frames.shape[0], (T, E, frames.shape[0])
            self._launch(self._fwd_idx, 'xa_nature_cnn_forward_indexed', ctypes.byref(self.net), ctypes.c_void_p(frames.data_ptr()),
                         int(frames.shape[0]), ctypes.c_void_p(idx.data_ptr()), T, E, self._stream(stream))
        else:
            self._launch(self._fwd, 'xa_nature_cnn_forward', ctypes.byref(self.net), ctypes.c_void_p(frames.data_ptr()), int(s2d),
                         self._stream(stream))
        ops._count(6)
        return actor, critic

    def backward(self, d_actor, d_critic, flat_grad, stream=None):
        """Output gradients (fp32, from the loss kernel) -> every element of `flat_grad` (the model's flat gradient buffer)."""
        assert self.has_backward and self._x1_in is not None, 'forward() first, on a plan built with a flat parameter buffer'
        assert d_actor.dtype == torch.float32 and d_actor.is_contiguous() and d_actor.numel() == self.batch * self.pack.n_actions
        assert d_critic.dtype == torch.float32 and d_critic.is_contiguous() and d_critic.numel() == self.batch
        assert flat_grad.dtype == torch.float32 and flat_grad.numel() >= self.net.n_grad
        self._launch(self._bwd, 'xa_nature_cnn_backward', ctypes.byref(self.net), ctypes.c_void_p(self._x1_in.data_ptr()),
                     ctypes.c_void_p(d_actor.data_ptr()), ctypes.c_void_p(d_critic.data_ptr()), ctypes.c_void_p(flat_grad.data_ptr()),
                     self._stream(stream))
        ops._count(9)

    def _launch(self, fn, name, *args):
        if self.device.index != torch.cuda.current_device():
            with torch.cuda.device(self.device):
                return _ffi.check(name, fn(*args))
        _ffi.check(name, fn(*args))
