"""Dense layers of the policy/value network on the tcgen05 GEMM (first slice).

`TcLinear` is a drop-in for torch.nn.Linear (fp32 master weights, same init contract) whose three products
run through `xa_gemm_bf16_tn` -- bf16 operands, fp32 accumulation in TMEM:

    forward   y  = x  W^T (+ b) (ReLU fused in the epilogue)        A = x [M,K],      B = W   [N,K]
    backward  dx = dy W                                            A = dy [M,N],     B = W^T [K,N]
              dW = dy^T x       `xa_gemm_bf16_atb`: dy [M,N] and x [M,K] as they are (MN-major UMMA operands)

bf16 copies of the weights are refreshed by `refresh()` after every optimiser step (the fused clip+Adam kernel
writes the fp32 master in place).  Convolutions still go through the framework's library path; the
implicit-GEMM convolutions are the next slice of this row.
"""
import torch

from .. import ops


class _TcLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu, w16, wt16):
        x2 = x.reshape(-1, x.shape[-1])
        x16 = x2 if (x2.dtype == torch.bfloat16 and x2.is_contiguous()) else ops.to_bf16(x2.float().contiguous() if x2.dtype not in (torch.float32, torch.bfloat16) else x2.contiguous())
        y = ops.gemm_bf16_tn(x16, w16, bias=bias, relu=relu, out_dtype=torch.float32)
        ctx.save_for_backward(x16, y if relu else None, wt16)
        ctx.relu, ctx.has_bias, ctx.x_shape, ctx.x_dtype = relu, bias is not None, x.shape, x.dtype
        return y.reshape(x.shape[:-1] + (weight.shape[0],))

    @staticmethod
    def backward(ctx, dy):
        x16, y, wt16 = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1]).float()
        if ctx.relu:
            dy2 = dy2 * (y > 0)
        dy2 = dy2.contiguous()
        n = dy2.shape[1]
        pad = (-n) % 8
        dy16 = ops.to_bf16(dy2) if pad == 0 else torch.nn.functional.pad(dy2, (0, pad)).to(torch.bfloat16)
        dx = ops.gemm_bf16_tn(dy16, wt16, out_dtype=torch.float32)                  # [M, K]
        if x16.shape[1] % 8 == 0:
            dw = ops.gemm_bf16_atb(dy16, x16)[:n]                                   # dy^T x from the natural layouts
        else:
            dw = ops.gemm_bf16_tn(ops.to_bf16(dy2, transpose=True), ops.to_bf16(x16, transpose=True), out_dtype=torch.float32)
        db = dy2.sum(0) if ctx.has_bias else None
        return dx.reshape(ctx.x_shape).to(ctx.x_dtype), dw, db, None, None, None


class TcLinear(torch.nn.Module):
    def __init__(self, in_features, out_features, bias=True, relu=False):
        super().__init__()
        assert in_features % 8 == 0, 'in_features must be a multiple of 8 (16-byte TMA row pitch)'
        self.in_features, self.out_features, self.relu = in_features, out_features, relu
        self.weight = torch.nn.Parameter(torch.empty(out_features, in_features))
        self.bias = torch.nn.Parameter(torch.zeros(out_features)) if bias else None
        torch.nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        self._w16 = self._wt16 = None

    def refresh(self):
        """Re-derive the bf16 operand copies from the fp32 master weights (call after each optimiser step)."""
        w = self.weight.detach()
        self._w16 = ops.to_bf16(w.contiguous())
        self._wt16 = ops.to_bf16(w.contiguous(), transpose=True)                    # [K, N'] with N' = N padded to 8
        return self

    def forward(self, x):
        if self._w16 is None or self._w16.device != self.weight.device:
            self.refresh()
        return _TcLinearFn.apply(x, self.weight, self.bias, self.relu, self._w16, self._wt16)
