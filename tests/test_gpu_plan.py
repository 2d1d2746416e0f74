"""The native two-call form of the policy/value network (csrc/nature_net.cu, heads.cu, grad_finalize.cu; agents/tc_plan.py):
heads kernels and the gradient finalisation against plain fp32 torch, the plan against the autograd form of the same
pipeline and against the bf16-emulating fp64 reference of test_gpu_conv.py."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _call(name, *args):
    from xagents_b200 import _ffi
    _ffi.check(name, getattr(_ffi.lib(), name)(*args))


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize('B,A', [(1, 6), (37, 6), (256, 4), (8192, 7)])
def test_heads_forward_vs_torch(B, A):
    torch.manual_seed(B)
    h = torch.randn(B, 512, device=DEV).relu().bfloat16()
    wh = torch.zeros(8, 512, device=DEV)
    wh[:A + 1] = torch.randn(A + 1, 512, device=DEV) * 0.05
    wh16 = wh.bfloat16()
    bh = torch.zeros(8, device=DEV)
    bh[:A + 1] = torch.randn(A + 1, device=DEV)
    actor = torch.full((B, A), float('nan'), device=DEV)
    critic = torch.full((B,), float('nan'), device=DEV)
    _call('xa_heads_forward_bf16', _p(h), _p(wh16), _p(bh), _p(actor), _p(critic), B, 512, A, _s())
    ref = h.float() @ wh16.float().t() + bh
    torch.testing.assert_close(actor, ref[:, :A], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(critic, ref[:, A], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('B,A', [(1, 6), (37, 6), (256, 4), (1000, 7)])
def test_heads_forward_from_split_k_partials_is_bit_identical_to_gemm_then_heads(B, A):
    """xa_gemm_bf16_tn_partial + xa_heads_forward_partial_bf16 (the FC layer's split-K partial sums are added, biased, rectified and
    rounded by the heads kernel: one launch less per rollout step) against xa_gemm_bf16_tn_ex (its own reduction pass) +
    xa_heads_forward_bf16: h, logits and values identical bit for bit."""
    from xagents_b200 import _ffi
    lib = _ffi.lib()
    torch.manual_seed(B)
    K, H = 3136, 512
    y3 = (torch.randn(B, K, device=DEV).relu() * 0.3).bfloat16()
    wf = (torch.randn(H, K, device=DEV) * 0.02).bfloat16()
    bf = torch.randn(H, device=DEV) * 0.1
    wh = torch.zeros(8, H, device=DEV)
    wh[:A + 1] = torch.randn(A + 1, H, device=DEV) * 0.05
    wh16 = wh.bfloat16()
    bh = torch.zeros(8, device=DEV)
    bh[:A + 1] = torch.randn(A + 1, device=DEV)
    ws_bytes = lib.xa_gemm_workspace_bytes(B, H, K)
    assert ws_bytes > 0, 'these batch sizes are split along K'
    out = {}
    for fused in (False, True):
        ws = torch.full((ws_bytes // 4,), float('nan'), device=DEV)
        h = torch.full((B, H), float('nan'), device=DEV).bfloat16()
        actor = torch.full((B, A), float('nan'), device=DEV)
        critic = torch.full((B,), float('nan'), device=DEV)
        if fused:
            splits = ctypes.c_int(0)
            _call('xa_gemm_bf16_tn_partial', _p(y3), _p(wf), B, H, K, _p(ws), ws_bytes, ctypes.byref(splits), _s())
            assert splits.value > 1
            _call('xa_heads_forward_partial_bf16', _p(ws), splits.value, _p(bf), _p(h), _p(wh16), _p(bh), _p(actor), _p(critic), B, H, A, _s())
        else:
            _call('xa_gemm_bf16_tn_ex', _p(y3), _p(wf), _p(h), _p(bf), B, H, K, H, 1, 1, None, H, 0, 0, _p(ws), ws_bytes, _s())
            _call('xa_heads_forward_bf16', _p(h), _p(wh16), _p(bh), _p(actor), _p(critic), B, H, A, _s())
        torch.cuda.synchronize()
        out[fused] = (h, actor, critic)
    for a, b in zip(out[False], out[True]):
        assert torch.isfinite(a.float()).all() and torch.equal(a, b)
    ref = (y3.float() @ wf.float().t() + bf).relu()
    torch.testing.assert_close(out[True][0].float(), ref, rtol=2e-2, atol=2e-2)
    # too small a workspace: nothing is launched and the caller is told to take the plain path
    splits = ctypes.c_int(7)
    _call('xa_gemm_bf16_tn_partial', _p(y3), _p(wf), B, H, K, _p(ws), 16, ctypes.byref(splits), _s())
    assert splits.value == 1


@pytest.mark.parametrize('B,A', [(1, 6), (37, 6), (100, 3), (8192, 6), (40001, 6)])
def test_heads_backward_vs_torch(B, A):
    from xagents_b200 import _ffi
    torch.manual_seed(B + 1)
    h = torch.randn(B, 512, device=DEV).relu().bfloat16()
    wh = torch.zeros(8, 512, device=DEV)
    wh[:A + 1] = torch.randn(A + 1, 512, device=DEV) * 0.05
    wh16 = wh.bfloat16()
    d_actor, d_critic = torch.randn(B, A, device=DEV) / B, torch.randn(B, device=DEV) / B
    blocks = _ffi.lib().xa_heads_backward_blocks(B)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    rows_per_cta = min(64, max(2, (-(-B // (2 * sms)) + 1) // 2 * 2))       # an even count that keeps every CTA resident (two per SM)
    assert blocks == -(-B // rows_per_cta)
    partial = torch.full((blocks, 10, 512), float('nan'), device=DEV)
    dh = torch.full((B, 512), float('nan'), device=DEV, dtype=torch.bfloat16)
    _call('xa_heads_backward_bf16', _p(d_actor), _p(d_critic), _p(h), _p(wh16), _p(dh), _p(partial), partial.numel(), B, 512, A, _s())
    d_out = torch.zeros(B, 8, device=DEV)
    d_out[:, :A], d_out[:, A] = d_actor, d_critic
    ref_dh = ((d_out @ wh16.float()) * (h > 0)).bfloat16()
    # one fma order against torch's: ties of the bf16 rounding may flip
    assert ((dh.float() - ref_dh.float()).abs() <= 1e-2 * ref_dh.float().abs() + 1e-9).all()
    assert (dh != ref_dh).float().mean() < 2e-2
    torch.testing.assert_close(partial[:, :8].sum(0), d_out.t() @ h.float(), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(partial[:, 8].sum(0), dh.float().sum(0), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(partial[:, 9, :8].sum(0), d_out.sum(0), rtol=1e-4, atol=1e-7)


def test_grad_finalize_adds_splits_in_order_and_permutes():
    from xagents_b200 import _ffi
    torch.manual_seed(3)
    # three segments: 1000 outputs x 37 splits (wide), 5000 outputs x 1 split through a permutation, 300 outputs x 2 splits;
    # 4 padding outputs without a source
    n = 1000 + 5000 + 300 + 4
    src = torch.randn(37 * 1000 + 5000 + 2 * 300, device=DEV)
    gmap = torch.full((n,), -1, dtype=torch.int32, device=DEV)
    gmap[:1000] = torch.randperm(1000, device=DEV).int()
    gmap[1000:6000] = 37000 + torch.randperm(5000, device=DEV).int()
    gmap[6000:6300] = 42000 + torch.arange(300, device=DEV).int()
    segs = (_ffi.GradSegment * 3)()
    for s, (begin, stride, splits, wide) in zip(segs, ((0, 1000, 37, 1), (1000, 0, 1, 0), (6000, 300, 2, 0))):
        s.dest_begin, s.split_stride, s.splits, s.wide = begin, stride, splits, wide
    grad = torch.full((n,), float('nan'), device=DEV)
    _call('xa_grad_finalize_f32', _p(src), _p(gmap), None, segs, 3, _p(grad), n, _s())
    a = src[:37000].view(37, 1000)[:, gmap[:1000].long()]
    # four lanes, each a contiguous quarter of the splits in order, then (q0 + q1) + (q2 + q3)
    quarters = []
    for q in range(4):
        acc = torch.zeros(1000, device=DEV)
        for k in range(q * 10, min(37, q * 10 + 10)):
            acc = acc + a[k]
        quarters.append(acc)
    assert torch.equal(grad[:1000], (quarters[0] + quarters[1]) + (quarters[2] + quarters[3]))
    assert torch.equal(grad[1000:6000], src[gmap[1000:6000].long()])
    b = src[42000:].view(2, 300)
    assert torch.equal(grad[6000:6300], b[0] + b[1])
    assert torch.equal(grad[6300:], torch.zeros(4, device=DEV))
    # the same sums stored through a destination permutation
    dest = torch.randperm(n, device=DEV).int()
    out = torch.full((n,), float('nan'), device=DEV)
    _call('xa_grad_finalize_f32', _p(src), _p(gmap), _p(dest), segs, 3, _p(out), n, _s())
    assert torch.equal(out[dest.long()], grad)
    # eight lanes per output (few outputs, hundreds of splits: the heads kernel's per-CTA blocks): contiguous eighths of the
    # splits in order, then ((p0 + p1) + (p2 + p3)) + ((p4 + p5) + (p6 + p7))
    src8 = torch.randn(300 * 100, device=DEV)
    map8 = torch.randperm(100, device=DEV).int()
    seg8 = (_ffi.GradSegment * 1)()
    seg8[0].dest_begin, seg8[0].split_stride, seg8[0].splits, seg8[0].wide = 0, 100, 300, 8
    got8 = torch.full((100,), float('nan'), device=DEV)
    _call('xa_grad_finalize_f32', _p(src8), _p(map8), None, seg8, 1, _p(got8), 100, _s())
    a8 = src8.view(300, 100)[:, map8.long()]
    parts = []
    for q in range(8):
        acc = torch.zeros(100, device=DEV)
        for k in range(q * 38, min(300, q * 38 + 38)):
            acc = acc + a8[k]
        parts.append(acc)
    assert torch.equal(got8, ((parts[0] + parts[1]) + (parts[2] + parts[3])) + ((parts[4] + parts[5]) + (parts[6] + parts[7])))


@pytest.mark.timeout(300)
@pytest.mark.parametrize('B,s2d_input', [(37, False), (512, True), (1000, False)])
def test_native_plan_matches_the_autograd_pipeline_and_the_emulated_reference(B, s2d_input):
    """TorchModel with the native plan against (a) the same network through torch autograd around the same kernels and
    (b) the fp64 reference that rounds to bf16 where the pipeline does."""
    from test_gpu_conv import _emulated_backward
    from xagents_b200 import ops
    from xagents_b200.agents import NatureCNN, NatureCnnTc, TorchModel
    torch.manual_seed(4)
    mods = [NatureCnnTc(4, 6).cuda() for _ in range(2)]
    with torch.no_grad():
        for p in mods[0].parameters():
            if p.dim() == 1:
                p.copy_(torch.randn_like(p) * 0.05)
        mods[0].actor.weight.mul_(30)
    mods[1].load_state_dict(mods[0].state_dict())
    ref = NatureCNN(4, 6).cuda()
    ref.load_state_dict(mods[0].state_dict())
    native, auto = TorchModel(mods[0], native_plan=True), TorchModel(mods[1], native_plan=False)
    x = torch.randint(0, 256, (B, 84, 84, 4), dtype=torch.uint8, device=DEV)
    d_actor, d_critic = torch.randn((B, 6), device=DEV) / B, torch.randn(B, device=DEV) / B
    frames = ops.space_to_depth_u8_bf16(x, 4) if s2d_input else x
    outs = []
    for m in (native, auto):
        a, c = m.forward(frames, training=True)
        outs.append((a.clone(), c.clone()))
        m.lr = 0.0                                                       # keep the weights: only the gradients are compared
        m.backward_and_step(d_actor, d_critic, grad_norm=None)
    torch.cuda.synchronize()
    assert native._native_plan and not auto._native_plan
    torch.testing.assert_close(outs[0][0], outs[1][0], rtol=2e-3, atol=2e-3)     # the heads product: fp32 fma order vs the GEMM's
    torch.testing.assert_close(outs[0][1], outs[1][1], rtol=2e-3, atol=2e-3)
    with torch.no_grad():
        emu = _emulated_backward(ref, x, d_actor, d_critic)
    for (name, p), q in zip(mods[0].named_parameters(), mods[1].parameters()):
        g, r, e = p.grad.flatten().double(), q.grad.flatten().double(), emu[name].flatten()
        rel_auto, rel_emu = float((g - r).norm() / r.norm()), float((g - e).norm() / e.norm())
        print(f'{name:18s} vs autograd pipeline rel-L2 {rel_auto:.5f} | vs bf16-emulated fp64 rel-L2 {rel_emu:.5f}')
        # the autograd pipeline rounds d_out to bf16 before the heads products; the plan keeps it in fp32
        assert rel_auto < 2e-2, f'{name}: {rel_auto:.5f} against the autograd pipeline'
        assert rel_emu < 2e-2, f'{name}: {rel_emu:.5f} against the bf16-emulated reference'
    assert torch.isfinite(native.flat_grad).all()
    # a second minibatch through the same plan: buffers are reused, nothing accumulates
    g1 = native.flat_grad.clone()
    native.forward(frames, training=True)
    native.backward_and_step(d_actor, d_critic, grad_norm=None)
    torch.cuda.synchronize()
    assert torch.equal(g1, native.flat_grad), 'the native backward is deterministic and overwrites its outputs'


@pytest.mark.timeout(300)
def test_ppo_train_step_through_the_native_plan_matches_the_autograd_pipeline():
    """Two PPO agents on the same replayed environments and seeds, one per network path: the same losses to bf16 accuracy."""
    import importlib.util
    import os

    import numpy as np
    from xagents_b200.agents import PPO, NatureCnnTc, TorchModel
    spec = importlib.util.spec_from_file_location('make_golden', os.path.join(os.path.dirname(__file__), 'golden', 'make_golden.py'))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    T, E, A = 8, 4, 6
    losses = []
    for native in (True, False):
        rng = np.random.default_rng(11)
        obs, rewards, dones, resets = mg._streams(rng, 4 * T, E, (84, 84, 4), True, 0.1)
        envs = [mg.ReplayEnv(obs[i], rewards[i], dones[i], resets[i], mg.Discrete(A)) for i in range(E)]
        torch.manual_seed(0)
        net = TorchModel(NatureCnnTc(4, A).cuda(), native_plan=native)
        agent = PPO(envs, net, n_steps=T, mini_batches=4, ppo_epochs=2, quiet=True, seed=5)
        agent.fit(max_steps=2 * T * E)
        torch.cuda.synchronize()
        assert net.step == 2 * 2 * 4 and torch.isfinite(net.flat_param).all()
        losses.append(torch.stack(agent.loss_history).cpu().numpy())
    assert np.isfinite(losses[0]).all()
    np.testing.assert_allclose(losses[0][:4], losses[1][:4], rtol=5e-2, atol=5e-3)


@pytest.mark.parametrize('B', [1, 3, 37, 256, 1000])
@pytest.mark.parametrize('N,out_s2d', [(32, True), (32, False), (64, False)])
def test_first_layer_from_uint8_frames_is_bit_identical_to_space_to_depth_then_convolution(B, N, out_s2d):
    from xagents_b200 import ops
    torch.manual_seed(B * 7 + N)
    x = torch.randint(0, 256, (B, 84, 84, 4), dtype=torch.uint8, device=DEV)
    w = (torch.randn(N, 256, device=DEV) * 0.05).bfloat16()
    bias = torch.randn(N, device=DEV) * 0.1
    want = ops.conv2d_nhwc_bf16(ops.space_to_depth_u8_bf16(x, 4), w, 2, 2, bias=bias, relu=True, out_s2d=out_s2d)
    got = ops.conv2d_u8_s2d_bf16(x, w, 2, 2, bias=bias, relu=True, out_s2d=out_s2d)
    x1 = torch.full((B, 21, 21, 64), float('nan'), dtype=torch.bfloat16, device=DEV)
    got2 = ops.conv2d_u8_s2d_bf16(x, w, 2, 2, bias=bias, relu=True, out_s2d=out_s2d, x_s2d_out=x1)
    torch.cuda.synchronize()
    assert got.shape == want.shape
    assert torch.equal(got.view(torch.int16), want.view(torch.int16)) and torch.equal(got2.view(torch.int16), want.view(torch.int16))
    assert torch.equal(x1.view(torch.int16), ops.space_to_depth_u8_bf16(x, 4).view(torch.int16)), 'the stored space-to-depth tensor'

    # extreme byte values in every position of a window
    for fill in (0, 255):
        xf = torch.full_like(x[:2], fill)
        assert torch.equal(ops.conv2d_u8_s2d_bf16(xf, w, 2, 2, bias=bias, relu=False).view(torch.int16),
                           ops.conv2d_nhwc_bf16(ops.space_to_depth_u8_bf16(xf, 4), w, 2, 2, bias=bias, relu=False).view(torch.int16))


@pytest.mark.parametrize('B,T,E', [(1, 0, 0), (37, 0, 0), (256, 8, 40), (1000, 16, 64), (2048, 128, 16)])
def test_first_layer_reads_the_minibatch_through_the_permutation(B, T, E):
    """xa_conv2d_u8_s2d_bf16_indexed: the tf.gather of the states (ppo/agent.py:139-155) folded into the first layer -- per-frame
    TMA bulk copies addressed through the ids -- is bit-identical to gathering the frames and running the layer on the copy:
    outputs and the stored space-to-depth tensor, plain row ids and env-major sample ids of a time-major rollout, duplicates."""
    from xagents_b200 import ops
    torch.manual_seed(B + 3)
    n_frames = T * E if T else 300
    store = torch.randint(0, 256, (n_frames, 84, 84, 4), dtype=torch.uint8, device=DEV)
    idx = torch.randint(0, n_frames, (B,), dtype=torch.int32, device=DEV)      # with repetitions
    idx[0], idx[-1] = n_frames - 1, 0
    rows = ((idx % T) * E + idx // T).long() if T else idx.long()             # base.py:559-564
    w = (torch.randn(32, 256, device=DEV) * 0.05).bfloat16()
    bias = torch.randn(32, device=DEV) * 0.1
    want_x1 = torch.empty((B, 21, 21, 64), dtype=torch.bfloat16, device=DEV)
    want = ops.conv2d_u8_s2d_bf16(store[rows].contiguous(), w, 2, 2, bias=bias, relu=True, out_s2d=True, x_s2d_out=want_x1)
    x1 = torch.full((B, 21, 21, 64), float('nan'), dtype=torch.bfloat16, device=DEV)
    got = ops.conv2d_u8_s2d_bf16(store, w, 2, 2, bias=bias, relu=True, out_s2d=True, x_s2d_out=x1, idx=idx,
                                 time_major=(T, E) if T else None)
    torch.cuda.synchronize()
    assert got.shape == want.shape and torch.equal(got.view(torch.int16), want.view(torch.int16))
    assert torch.equal(x1.view(torch.int16), want_x1.view(torch.int16))
    if T:
        with pytest.raises(Exception, match='not the number of stored frames'):
            ops.conv2d_u8_s2d_bf16(store, w, 2, 2, idx=idx, time_major=(T + 1, E))
