"""The Nature CNN, forward AND backward, on the tcgen05 kernels.

`NatureCnnTc` keeps fp32 master weights in the torch layouts of `NatureCNN` (so checkpoints, the flat-buffer
optimiser and the gradient all-reduce are unchanged) and runs every contraction of the network through
`xa_conv2d_nhwc_bf16` / `xa_gemm_bf16_tn` (bf16 operands, fp32 accumulation in TMEM):

  forward    frames -s2d-> conv1 -> conv2 -> conv3 -> FC512 -> heads                     (tc_conv.py's pipeline)
  backward   heads/FC: dgrad GEMMs with the ReLU derivative in the epilogue, wgrad = `xa_gemm_bf16_atb` (dY^T X from the
             natural row-major tensors, MN-major UMMA operands)
             conv3, conv2: data gradient = the same convolution kernel on dY with full zero padding and flipped
             weights (+ ReLU derivative of the layer below in the epilogue)
             all convs: weight + bias gradient = `xa_conv_wgrad_nhwc_bf16`: a shifted-window GEMM over the NATURAL
             NHWC tensors (MN-major UMMA operands: no transposes, no im2col matrix), split over the SMs along the
             pixel axis.  For that every dY is kept on the zero-bordered pixel grid of its layer's INPUT, which the
             producing kernels write directly (column-group map of the FC data-gradient GEMM, output grid / unpack
             options of the data-gradient convolutions)
  The input layer needs no data gradient.

Strided layers live in space-to-depth form (8x8/4 -> 2x2/1 over 21x21x64, 4x4/2 -> 2x2/1 over 10x10x128); the
layout maps between torch's [N, C, KH, KW] and the kernels' [N, (kh, kw, c)] are applied to the (tiny) weight and
weight-gradient tensors only.
"""
import torch

from .. import ops
from .models import NatureCNN
from .tc_operands import OperandPack


def _s2d_kernel_inverse(g, n, c, kh, kw, s):
    """[N, (kh', kw', dy, dx, c)] -> torch [N, C, KH, KW]."""
    g = g.reshape(n, kh // s, kw // s, s, s, c).permute(0, 1, 3, 2, 4, 5).reshape(n, kh, kw, c)
    return g.permute(0, 3, 1, 2).contiguous()


class _NatureCnnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, frames, op, w1, b1, w2, b2, w3, b3, wf, bf_, wa, ba, wc, bc):
        # uint8 frames, or the output of ops.gather_s2d_u8_bf16 (minibatch gather fused with /255 + space-to-depth)
        x1 = frames if frames.dtype == torch.bfloat16 else ops.space_to_depth_u8_bf16(frames.contiguous(), 4)   # [B,21,21,64]
        x2 = ops.conv2d_nhwc_bf16(x1, op.w1, 2, 2, bias=op.b1, relu=True, out_s2d=True)           # [B,10,10,128]
        x3 = ops.conv2d_nhwc_bf16(x2, op.w2, 2, 2, bias=op.b2, relu=True)                         # [B,9,9,64]
        y3 = ops.conv2d_nhwc_bf16(x3, op.w3, 3, 3, bias=op.b3, relu=True)                         # [B,7,7,64]
        h = ops.gemm_bf16_tn(y3.view(y3.shape[0], -1), op.wf, bias=op.bf_, relu=True, out_dtype=torch.bfloat16)
        out = ops.gemm_bf16_tn(h, op.wh, bias=op.bh, out_dtype=torch.float32)                     # [B,8]
        ctx.op = op
        ctx.save_for_backward(x1, x2, x3, y3, h)
        return out[:, :op.n_actions].contiguous(), out[:, op.n_actions].contiguous()

    @staticmethod
    def backward(ctx, d_actor, d_critic):
        op = ctx.op
        x1, x2, x3, y3, h = ctx.saved_tensors
        B, A = h.shape[0], op.n_actions
        d_out = torch.zeros((B, op.wh.shape[0]), dtype=torch.float32, device=h.device)
        d_out[:, :A] = d_actor
        d_out[:, A] = d_critic.reshape(-1)
        # heads
        d_out16 = ops.to_bf16(d_out)
        d_wh = ops.gemm_bf16_atb(d_out16, h)                                                                 # [8,512] = d_out^T h
        d_bh = d_out.sum(0)
        dh = ops.gemm_bf16_tn(d_out16, op.wh_t, relu_mask=h, out_dtype=torch.bfloat16)                       # [B,512]
        # FC512
        y3f = y3.view(B, -1)
        d_wf = ops.gemm_bf16_atb(dh, y3f)                                                                    # [512,3136] = dh^T y3
        d_bf = dh.sum(0, dtype=torch.float32)
        g3, g2, g1 = op.grids(B)
        ops.gemm_bf16_tn(dh, op.wf_t, relu_mask=y3f, out=g3.view(B, -1), col_group=(7 * 64, 9 * 64))         # 7x7 on the 9x9 grid
        # convolutions: dW, db by the shifted-window GEMM over natural NHWC tensors; dX by the padded / flipped convolution,
        # written straight onto the next layer's zero-bordered grid
        d_w3, d_b3 = ops.conv_wgrad_nhwc_bf16(x3, g3, 3, 3)                                                  # [64,576]
        ops.conv2d_nhwc_bf16(g3, op.w3_flip, 3, 3, pad=(2, 2), out_hw=(9, 9), relu_mask=x3, out=g2, zero_border=True)          # 9x9 on the 10x10 grid
        d_w2, d_b2 = ops.conv_wgrad_nhwc_bf16(x2, g2, 2, 2)                                                  # [64,512]
        ops.conv2d_nhwc_bf16(g2, op.w2_flip, 2, 2, pad=(1, 1), out_hw=(10, 10), relu_mask=x2, out=g1,        # [B,10,10,128] unpacked
                             unpack_s2d=True, zero_border=True)                                                                # to 20x20 on the 21x21 grid
        d_w1, d_b1 = ops.conv_wgrad_nhwc_bf16(x1, g1, 2, 2)                                                  # [32,256]
        # back to torch layouts
        g_w1 = _s2d_kernel_inverse(d_w1, 32, 4, 8, 8, 4)
        g_w2 = _s2d_kernel_inverse(d_w2, 64, 32, 4, 4, 2)
        g_w3 = d_w3.reshape(64, 3, 3, 64).permute(0, 3, 1, 2).contiguous()
        g_wf = d_wf.reshape(512, 7, 7, 64).permute(0, 3, 1, 2).reshape(512, -1).contiguous()
        return (None, None, g_w1, d_b1, g_w2, d_b2, g_w3, d_b3, g_wf, d_bf, d_wh[:A].contiguous(), d_bh[:A].contiguous(),
                d_wh[A:A + 1].contiguous(), d_bh[A:A + 1].contiguous())


class NatureCnnTc(NatureCNN):
    """NatureCNN whose forward and backward run on the tcgen05 kernels.  Takes uint8 NHWC frames directly."""
    takes_uint8 = True          # ... or the bf16 space-to-depth tensor of ops.gather_s2d_u8_bf16

    def __init__(self, in_channels=4, n_actions=6):
        assert in_channels == 4, 'the space-to-depth layouts are built for 84x84x4 frames'
        super().__init__(in_channels, n_actions)
        self._op = None
        self._plans = {}

    def refresh(self):
        """bf16 operand layouts of the current weights: two launches (index-map gather + cast, tc_operands.py)."""
        if self._op is None:
            self._op = OperandPack(self)
        else:
            self._op.refresh()
        return self

    def plan(self, batch, flat_param=None, n_grad=None):
        """The native two-call form of this network for one batch size (tc_plan.NaturePlan).  With `flat_param` (the flat fp32
        buffer the parameters are views of, e.g. TorchModel.flat_param) the plan also back-propagates, writing straight into a
        flat gradient buffer of `n_grad` elements laid out like it."""
        if self._op is None:
            self.refresh()
        key = (int(batch), None if flat_param is None else flat_param.data_ptr(), n_grad)
        plan = self._plans.get(key)
        if plan is None:
            from .tc_plan import NaturePlan
            if len(self._plans) >= 4:                              # minibatch / rollout / bootstrap sizes; do not hoard activations
                self._plans.pop(next(iter(self._plans)))
            plan = self._plans[key] = NaturePlan(self._op, self._op._named, flat_param, batch, n_grad=n_grad, backward=flat_param is not None)
        return plan

    def forward(self, frames_u8):
        if self._op is None:
            self.refresh()
        c1, c2, c3 = [x for x in self.trunk if isinstance(x, torch.nn.Conv2d)]
        fc = [x for x in self.trunk if isinstance(x, torch.nn.Linear)][0]
        return _NatureCnnFn.apply(frames_u8, self._op, c1.weight, c1.bias, c2.weight, c2.bias, c3.weight, c3.bias, fc.weight, fc.bias,
                                  self.actor.weight, self.actor.bias, self.critic.weight, self.critic.bias)
