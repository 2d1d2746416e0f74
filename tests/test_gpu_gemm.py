"""tcgen05 GEMM (dense layers of the policy/value network) against a plain torch fp32 reference of the
same op on the same bf16-rounded operands.  fp32 accumulation on both sides: tolerance 1e-5 of max|C|
plus bf16 rounding when the output is bf16."""
import numpy as np
import pytest
import torch

from xagents_b200 import ops

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.timeout(120)
@pytest.mark.parametrize('m,n,k', [(128, 128, 64), (128, 128, 512), (256, 512, 3136), (8192, 512, 3136), (8192, 7, 512),
                                   (300, 70, 136), (1, 16, 8), (129, 130, 72), (4096, 64, 576)])
@pytest.mark.parametrize('bias,relu,out_bf16', [(False, False, False), (True, True, False), (True, False, True)])
def test_gemm_bf16_tn_vs_torch_fp32(m, n, k, bias, relu, out_bf16):
    g = torch.Generator(device=DEV)
    g.manual_seed(m * 131 + n * 7 + k)
    a = torch.randn((m, k), device=DEV, generator=g).to(torch.bfloat16)
    b = (torch.randn((n, k), device=DEV, generator=g) / k ** 0.5).to(torch.bfloat16)
    bv = torch.randn(n, device=DEV, generator=g) if bias else None
    want = a.float().double() @ b.float().double().t()
    if bias:
        want = want + bv.double()
    if relu:
        want = want.clamp_min(0)
    got = ops.gemm_bf16_tn(a, b, bias=bv, relu=relu, out_dtype=torch.bfloat16 if out_bf16 else torch.float32)
    torch.cuda.synchronize()
    assert got.shape == (m, n) and got.dtype == (torch.bfloat16 if out_bf16 else torch.float32)
    scale = float(want.abs().max())
    err = float((got.double() - want).abs().max())
    tol = (4e-3 if out_bf16 else 1e-5) * scale
    assert err <= tol, f'max abs err {err:.3e} > {tol:.3e} (scale {scale:.3e})'


@pytest.mark.timeout(60)
def test_gemm_argument_errors():
    from xagents_b200._ffi import XAError
    a = torch.zeros((8, 12), dtype=torch.bfloat16, device=DEV)
    with pytest.raises(XAError):                                   # K % 8 != 0: TMA needs a 16-byte row pitch
        ops.gemm_bf16_tn(a, a)
    with pytest.raises(ValueError):
        ops.gemm_bf16_tn(torch.zeros((8, 16), dtype=torch.bfloat16, device=DEV), torch.zeros((8, 8), dtype=torch.bfloat16, device=DEV))
    with pytest.raises(TypeError):
        ops.gemm_bf16_tn(torch.zeros((8, 16), device=DEV), torch.zeros((8, 16), device=DEV))
