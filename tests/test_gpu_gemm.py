"""tcgen05 GEMM (dense layers of the policy/value network) against a plain torch fp32 reference of the
same op on the same bf16-rounded operands.  fp32 accumulation on both sides: tolerance 1e-5 of max|C|
plus bf16 rounding when the output is bf16."""
import pytest
import torch

from xagents_b200 import ops

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.timeout(120)
@pytest.mark.parametrize('m,n,k', [(128, 128, 64), (128, 128, 512), (256, 512, 3136), (8192, 512, 3136), (8192, 7, 512),
                                   (300, 70, 136), (1, 16, 8), (129, 130, 72), (4096, 64, 576),
                                   (1100, 300, 200), (2048, 3136, 512), (1280, 256, 64), (600, 520, 72)])   # CTA pairs (m >= 512, n >= 256): ragged M / N, one K block
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


@pytest.mark.timeout(120)
@pytest.mark.parametrize('m,k,n,relu', [(8192, 3136, 512, True), (1000, 512, 6, False), (37, 512, 1, False)])
def test_tc_linear_forward_backward_vs_torch(m, k, n, relu):
    """TcLinear (three tcgen05 GEMMs + transposes) against torch.nn.Linear in fp32 on the same bf16-rounded
    weights and inputs: forward, dx, dW, db."""
    from xagents_b200.agents.tc_dense import TcLinear
    torch.manual_seed(m + n)
    lin = TcLinear(k, n, relu=relu).cuda()
    with torch.no_grad():
        lin.weight.copy_(lin.weight.to(torch.bfloat16).float())          # make the bf16 operand copy exact
        lin.bias.copy_(torch.randn(n, device=DEV) * 0.1)
    lin.refresh()
    x = torch.randn((m, k), device=DEV).to(torch.bfloat16).float().requires_grad_(True)
    dy = torch.randn((m, n), device=DEV).to(torch.bfloat16).float()
    y = lin(x)
    y.backward(dy)
    xr = x.detach().double().requires_grad_(True)
    wr = lin.weight.detach().double().requires_grad_(True)
    br = lin.bias.detach().double().requires_grad_(True)
    yr = xr @ wr.t() + br
    if relu:
        yr = yr.clamp_min(0)
    yr.backward(dy.double())

    def near(got, want, rel):
        scale = float(want.abs().max())
        assert float((got.double() - want).abs().max()) <= rel * scale

    near(y, yr.detach(), 1e-5)
    near(lin.bias.grad, br.grad, 1e-5)
    # dy is re-rounded to bf16 only through the ReLU mask (exact), so the backward products are fp32-accumulate exact
    near(x.grad, xr.grad, 2e-5)
    near(lin.weight.grad, wr.grad, 2e-5)


@pytest.mark.timeout(180)
def test_nature_cnn_with_tensor_core_dense_trains():
    from xagents_b200.agents import NatureCNN, TorchModel
    torch.manual_seed(0)
    net = TorchModel(NatureCNN(4, 6, tensor_core_dense=True).cuda())
    ref = NatureCNN(4, 6).cuda()
    ref.load_state_dict({k: v.clone() for k, v in net.module.state_dict().items()}, strict=True)
    x = torch.randint(0, 256, (256, 84, 84, 4), dtype=torch.uint8, device=DEV)
    actor, critic = net.forward(x, training=True)
    ra, rc = ref(x.float() / 255.0)
    assert float((actor - ra).abs().max()) <= 2e-2 * float(ra.abs().max()) + 1e-3      # bf16 operands vs fp32
    assert float((critic - rc.reshape(-1)).abs().max()) <= 2e-2 * float(rc.abs().max()) + 1e-3
    before = net.flat_param.clone()
    net.backward_and_step(torch.randn_like(actor) / 256, torch.randn_like(critic) / 256, grad_norm=0.5)
    torch.cuda.synchronize()
    assert torch.isfinite(net.flat_param).all() and not torch.equal(before, net.flat_param)


@pytest.mark.timeout(120)
@pytest.mark.parametrize('m,n,k', [(512, 3136, 8192), (64, 576, 401408), (32, 256, 40000), (8, 512, 8192)])
def test_gemm_split_k_matches_unsplit_and_reference(m, n, k):
    """Weight-gradient shapes: few output tiles, long K -> split along K over the SMs + deterministic reduction."""
    from xagents_b200 import _ffi
    g = torch.Generator(device=DEV)
    g.manual_seed(k)
    a = torch.randn((m, k), device=DEV, generator=g).to(torch.bfloat16)
    b = (torch.randn((n, k), device=DEV, generator=g) / k ** 0.5).to(torch.bfloat16)
    assert _ffi.lib().xa_gemm_workspace_bytes(m, n, k) > 0
    split = ops.gemm_bf16_tn(a, b)
    again = ops.gemm_bf16_tn(a, b)
    plain = ops.gemm_bf16_tn(a, b, split_k=False)
    torch.cuda.synchronize()
    assert torch.equal(split, again)                               # deterministic
    want = a.double() @ b.double().t()
    scale = float(want.abs().max())
    # fp32 accumulation in TMEM over thousands of products is good to ~5e-5 of the result scale
    assert float((split.double() - want).abs().max()) <= 5e-5 * scale
    # one TMEM accumulator over 400k products loses low bits (tensor-core accumulation is not a chain of IEEE adds):
    # another reason the long-K products are split
    assert float((plain.double() - want).abs().max()) <= (5e-5 if k < 10000 else 2e-3) * scale


@pytest.mark.timeout(60)
def test_gemm_relu_mask_epilogue():
    a = torch.randn((300, 512), device=DEV).to(torch.bfloat16)
    b = torch.randn((200, 512), device=DEV).to(torch.bfloat16)
    mask = torch.randn((300, 200), device=DEV).to(torch.bfloat16)
    got = ops.gemm_bf16_tn(a, b, relu_mask=mask)
    want = (a.double() @ b.double().t()) * (mask > 0)
    assert float((got.double() - want).abs().max()) <= 1e-5 * float(want.abs().max())


@pytest.mark.timeout(120)
@pytest.mark.parametrize('k,m,n', [(64, 8, 8), (200, 128, 64), (8192, 512, 3136), (8192, 8, 512), (1000, 72, 136), (37, 512, 8)])
def test_gemm_atb_from_row_major_operands(k, m, n):
    """C = A^T B with both operands row-major [K, *] (MN-major UMMA descriptors) against fp64."""
    torch.manual_seed(k + m)
    a = torch.randn(k, m, device=DEV).to(torch.bfloat16)
    b = torch.randn(k, n, device=DEV).to(torch.bfloat16)
    got = ops.gemm_bf16_atb(a, b)
    want = a.double().t() @ b.double()
    assert got.shape == (m, n)
    assert float((got.double() - want).abs().max()) <= 2e-5 * float(want.abs().max()) * max(1.0, (k / 1000) ** 0.5)
    one = ops.gemm_bf16_atb(a, b, split_k=False)
    assert float((one.double() - want).abs().max()) <= 2e-3 * float(want.abs().max())


@pytest.mark.parametrize('m,n,k,group', [(8192, 3136, 512, (448, 576)), (1000, 3136, 512, (448, 576)), (300, 64, 72, None), (129, 288, 200, (96, 128))])
def test_gemm_with_the_relu_mask_as_bits_equals_the_bf16_mask_form(m, n, k, group):
    """xa_gemm_bf16_tn_maskbits (mask = one bit per element, epilogue through TMA stores of 32 x 32 boxes) against xa_gemm_bf16_tn_ex with
    the bf16 activation as the mask (shared-memory transpose + per-lane stores): identical outputs, untouched bytes stay untouched
    (the grouped-column form writes a [m, h*w*c] gradient onto a zero-bordered grid)."""
    import ctypes
    from xagents_b200 import _ffi
    lib = _ffi.lib()
    g = torch.Generator(device=DEV)
    g.manual_seed(m + n)
    a = (torch.randn(m, k, device=DEV, generator=g) * 0.5).bfloat16()
    b = (torch.randn(n, k, device=DEV, generator=g) * 0.1).bfloat16()
    act = torch.randn(m, n, device=DEV, generator=g).relu().bfloat16()          # the layer's output: about half of it zero
    bits = (act > 0).view(m, n // 32, 32).to(torch.int64)
    words = (bits << torch.arange(32, device=DEV)).sum(-1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32).contiguous()
    cg, cp = group if group is not None else (0, 0)
    ldc = (n // cg) * cp + 64 if group is not None else n + 8
    outs = []
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    for form in ('bf16', 'bits'):
        c = torch.full((m, ldc), 7.0, device=DEV).bfloat16()
        if form == 'bits':
            _ffi.check('maskbits', lib.xa_gemm_bf16_tn_maskbits(p(a), p(b), p(c), m, n, k, ldc, p(words), n, cg, cp, s))
        else:
            _ffi.check('ex', lib.xa_gemm_bf16_tn_ex(p(a), p(b), p(c), None, m, n, k, ldc, 1, 0, p(act), n, cg, cp, None, 0, s))
        torch.cuda.synchronize()
        outs.append(c)
    assert torch.equal(outs[0], outs[1])
    ref = (a.float() @ b.float().t()) * (act > 0)
    got = outs[1].float()
    if group is not None:
        got = got[:, :(n // cg) * cp].view(m, n // cg, cp)[:, :, :cg].reshape(m, n)
        assert bool((outs[1].float().view(m, -1)[:, :(n // cg) * cp].view(m, n // cg, cp)[:, :, cg:] == 7.0).all())   # the border columns
    else:
        got = got[:, :n]
    torch.testing.assert_close(got, ref, rtol=2e-2, atol=2e-2)
    assert bool((outs[1][:, -8:].float() == 7.0).all())
