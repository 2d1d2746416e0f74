"""Implicit-GEMM convolution on tcgen05 against a plain torch fp32 reference of the same op on the same
bf16-rounded operands (fp32 accumulation on both sides; the output is rounded to bf16 once)."""
import pytest
import torch

from xagents_b200 import ops

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def torch_conv_nhwc(x, w, kh, kw, bias, relu):
    """x [B,H,W,C], w [N, kh*kw*C] (K ordered kh,kw,c) -> [B,OH,OW,N] in fp64."""
    n, c = w.shape[0], x.shape[-1]
    wt = w.double().reshape(n, kh, kw, c).permute(0, 3, 1, 2)
    y = torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), wt, bias.double() if bias is not None else None)
    y = y.permute(0, 2, 3, 1)
    return y.clamp_min(0) if relu else y


@pytest.mark.timeout(180)
@pytest.mark.parametrize('B,H,W,C,kh,kw,N,s2d', [(7, 21, 21, 64, 2, 2, 32, True), (5, 10, 10, 128, 2, 2, 64, False),
                                                 (9, 9, 9, 64, 3, 3, 64, False), (300, 21, 21, 64, 2, 2, 32, True),
                                                 (1000, 10, 10, 128, 2, 2, 64, False), (1001, 9, 9, 64, 3, 3, 64, False),
                                                 (3, 12, 17, 64, 1, 3, 96, False)])
@pytest.mark.parametrize('bias,relu', [(True, True), (False, False)])
def test_conv2d_nhwc_vs_torch(B, H, W, C, kh, kw, N, s2d, bias, relu):
    g = torch.Generator(device=DEV)
    g.manual_seed(B + H * 7 + N)
    x = torch.randn((B, H, W, C), device=DEV, generator=g).to(torch.bfloat16)
    w = (torch.randn((N, kh * kw * C), device=DEV, generator=g) / (kh * kw * C) ** 0.5).to(torch.bfloat16)
    bv = torch.randn(N, device=DEV, generator=g) if bias else None
    want = torch_conv_nhwc(x.float(), w.float(), kh, kw, bv, relu)
    got = ops.conv2d_nhwc_bf16(x, w, kh, kw, bias=bv, relu=relu, out_s2d=s2d)
    torch.cuda.synchronize()
    OH, OW = H - kh + 1, W - kw + 1
    if s2d:                                      # undo the space-to-depth layout of the output
        got = got.reshape(B, OH // 2, OW // 2, 2, 2, N).permute(0, 1, 3, 2, 4, 5).reshape(B, OH, OW, N)
    assert got.shape == want.shape
    scale = float(want.abs().max())
    err = float((got.double() - want).abs().max())
    assert err <= 4e-3 * scale, f'max abs err {err:.3e} vs scale {scale:.3e}'      # one bf16 rounding of the output


def flip_for_dgrad(w, kh, kw, c):
    """forward weights [N, kh*kw*C] -> data-gradient weights [C, kh*kw*N]: W'[c, kh', kw', n] = W[n, KH-1-kh', KW-1-kw', c]."""
    n = w.shape[0]
    return w.reshape(n, kh, kw, c).flip(1, 2).permute(3, 1, 2, 0).reshape(c, kh * kw * n).contiguous()


@pytest.mark.timeout(180)
@pytest.mark.parametrize('B,H,W,C,kh,kw,N', [(5, 9, 9, 64, 3, 3, 64), (700, 9, 9, 64, 3, 3, 64), (6, 10, 10, 128, 2, 2, 64),
                                             (513, 10, 10, 128, 2, 2, 64)])
@pytest.mark.parametrize('masked', [False, True])
def test_conv_data_gradient_vs_autograd(B, H, W, C, kh, kw, N, masked):
    """dX of a stride-1 convolution = the same kernel on dY with full zero padding and flipped weights; the ReLU
    derivative of the layer below rides in the epilogue."""
    g = torch.Generator(device=DEV)
    g.manual_seed(B + C)
    x = torch.randn((B, H, W, C), device=DEV, generator=g).to(torch.bfloat16)
    w = (torch.randn((N, kh * kw * C), device=DEV, generator=g) / (kh * kw * C) ** 0.5).to(torch.bfloat16)
    OH, OW = H - kh + 1, W - kw + 1
    dy = torch.randn((B, OH, OW, N), device=DEV, generator=g).to(torch.bfloat16)
    xr = x.double().requires_grad_(True)
    yr = torch_conv_nhwc(torch.relu(xr) if masked else xr, w.float(), kh, kw, None, False)
    yr.backward(dy.double())
    got = ops.conv2d_nhwc_bf16(dy, flip_for_dgrad(w, kh, kw, C), kh, kw, pad=(kh - 1, kw - 1), relu_mask=x if masked else None)
    torch.cuda.synchronize()
    assert got.shape == (B, H, W, C)
    scale = float(xr.grad.abs().max())
    assert float((got.double() - xr.grad).abs().max()) <= 4e-3 * scale


@pytest.mark.timeout(60)
def test_space_to_depth_u8():
    x = torch.randint(0, 256, (5, 84, 84, 4), dtype=torch.uint8, device=DEV)
    got = ops.space_to_depth_u8_bf16(x, 4)
    want = (x.float() / 255.0).reshape(5, 21, 4, 21, 4, 4).permute(0, 1, 3, 2, 4, 5).reshape(5, 21, 21, 64).to(torch.bfloat16)
    assert torch.equal(got, want)
    # every byte value: the kernel's x * (1/255) must round to the same bf16 as the reference's true division (base.py:505-506)
    ramp = (torch.arange(84 * 84 * 4, device=DEV) % 256).to(torch.uint8).reshape(1, 84, 84, 4)
    assert torch.equal(ops.space_to_depth_u8_bf16(ramp, 4).reshape(-1).sort().values,
                       (ramp.float() / 255.0).to(torch.bfloat16).reshape(-1).sort().values)
    got = ops.space_to_depth_u8_bf16(x, 2, scale_255=False)
    want = x.float().reshape(5, 42, 2, 42, 2, 4).permute(0, 1, 3, 2, 4, 5).reshape(5, 42, 42, 16).to(torch.bfloat16)
    assert torch.equal(got, want)


@pytest.mark.timeout(180)
@pytest.mark.parametrize('B', [3, 256, 1030])
def test_nature_cnn_forward_on_tensor_cores_vs_torch(B):
    from xagents_b200.agents import NatureCNN
    from xagents_b200.agents.tc_conv import NatureCnnTcForward
    torch.manual_seed(1)
    net = NatureCNN(4, 6).cuda()
    with torch.no_grad():                          # non-trivial biases and a head that is not ~0
        for p in net.parameters():
            if p.dim() == 1:
                p.copy_(torch.randn_like(p) * 0.1)
        net.actor.weight.mul_(50)
    tc = NatureCnnTcForward(net)
    x = torch.randint(0, 256, (B, 84, 84, 4), dtype=torch.uint8, device=DEV)
    actor, critic = tc(x)
    rb = lambda t: t.to(torch.bfloat16).float()                      # the same roundings the bf16 pipeline makes
    with torch.no_grad():
        ra, rc = net(x.float() / 255.0)                               # plain fp32 network
        convs = [m for m in net.trunk if isinstance(m, torch.nn.Conv2d)]
        fc = [m for m in net.trunk if isinstance(m, torch.nn.Linear)][0]
        h = rb(x.float() / 255.0).permute(0, 3, 1, 2)
        for conv in convs:
            h = rb(torch.relu(torch.nn.functional.conv2d(h, rb(conv.weight), conv.bias, conv.stride)))
        h = h.permute(0, 2, 3, 1).reshape(B, -1)                      # NHWC flatten
        wf = fc.weight.reshape(512, 64, 7, 7).permute(0, 2, 3, 1).reshape(512, -1)
        h = rb(torch.relu(h @ rb(wf).t() + fc.bias))
        ea = h @ rb(net.actor.weight).t() + net.actor.bias
        ec = (h @ rb(net.critic.weight).t() + net.critic.bias).reshape(-1)
    torch.cuda.synchronize()
    assert actor.shape == (B, 6) and critic.shape == (B,)
    for got, emu, ref in ((actor, ea, ra), (critic, ec, rc.reshape(-1))):
        scale = float(ref.abs().max())
        assert float((got - emu).abs().max()) <= 2e-2 * scale          # same roundings; summation order can flip a bf16 tie
        assert float((got - ref).abs().max()) <= 8e-2 * scale          # five bf16 layers against the fp32 network


@pytest.mark.timeout(180)
def test_agent_rollout_uses_tensor_core_inference():
    """PPO.fit with rollout-time policy evaluation on the tcgen05 pipeline and training through torch autograd."""
    import importlib.util
    import os

    import numpy as np
    from xagents_b200.agents import PPO, NatureCNN, TorchModel
    spec = importlib.util.spec_from_file_location('make_golden', os.path.join(os.path.dirname(__file__), 'golden', 'make_golden.py'))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    T, E, A = 8, 4, 6
    rng = np.random.default_rng(7)
    obs, rewards, dones, resets = mg._streams(rng, 4 * T, E, (84, 84, 4), True, 0.1)
    envs = [mg.ReplayEnv(obs[i], rewards[i], dones[i], resets[i], mg.Discrete(A)) for i in range(E)]
    torch.manual_seed(0)
    net = TorchModel(NatureCNN(4, A).cuda(), tensor_core_inference=True)
    before = ops.launch_count()
    agent = PPO(envs, net, n_steps=T, mini_batches=4, ppo_epochs=2, quiet=True, seed=3)
    agent.fit(max_steps=2 * T * E)
    torch.cuda.synchronize()
    assert agent.steps == 2 * T * E and torch.isfinite(net.flat_param).all()
    assert ops.launch_count() - before > 2 * (T + 1) * 6            # 6 tensor-core launches per rollout forward
    # the bf16 inference path and the fp32 training path agree on the rollout's values
    x = torch.as_tensor(obs[:, 0]).cuda()
    _, v_tc = net.forward(x, training=False)
    net._tc_forward, keep = None, net._tc_forward
    _, v_ref = net.forward(x, training=False)
    net._tc_forward = keep
    assert float((v_tc - v_ref).abs().max()) <= 8e-2 * float(v_ref.abs().max()) + 1e-3


def _emulated_backward(net, x_u8, d_actor, d_critic):
    """fp64 forward+backward of the Nature CNN with bf16 roundings at exactly the points where the tensor-core
    pipeline rounds (weights, every activation, every back-propagated gradient tensor)."""
    rb = lambda t: t.to(torch.bfloat16).double()
    F = torch.nn.functional
    convs = [m for m in net.trunk if isinstance(m, torch.nn.Conv2d)]
    fc = [m for m in net.trunk if isinstance(m, torch.nn.Linear)][0]
    B = x_u8.shape[0]
    w = [rb(c.weight) for c in convs]
    acts = [rb(x_u8.float() / 255.0).permute(0, 3, 1, 2)]
    for c, wi in zip(convs, w):
        acts.append(rb(torch.relu(F.conv2d(acts[-1], wi, c.bias.double(), c.stride))))
    y3 = acts[-1].permute(0, 2, 3, 1).reshape(B, -1)                                  # NHWC flatten
    wf = rb(fc.weight.reshape(512, 64, 7, 7).permute(0, 2, 3, 1).reshape(512, -1))
    h = rb(torch.relu(y3 @ wf.t() + fc.bias.double()))
    A = net.actor.weight.shape[0]
    wh = torch.cat([rb(net.actor.weight), rb(net.critic.weight)])                      # [A+1, 512]
    d_out = torch.cat([d_actor.double(), d_critic.double().reshape(-1, 1)], 1)
    d_out16 = rb(d_out)
    g = {}
    d_wh = d_out16.t() @ h
    g['actor.weight'], g['critic.weight'] = d_wh[:A], d_wh[A:]
    g['actor.bias'], g['critic.bias'] = d_out[:, :A].sum(0), d_out[:, A:].sum(0)
    dh = rb((d_out16 @ wh) * (h > 0))
    d_wf = dh.t() @ y3
    g['trunk.7.weight'] = d_wf.reshape(512, 7, 7, 64).permute(0, 3, 1, 2).reshape(512, -1)
    g['trunk.7.bias'] = dh.sum(0)
    dy = rb((dh @ wf) * (y3 > 0)).reshape(B, 7, 7, 64).permute(0, 3, 1, 2)            # NCHW
    for li in (2, 1, 0):
        c = convs[li]
        g[f'trunk.{2 * li}.weight'] = torch.nn.grad.conv2d_weight(acts[li], w[li].shape, dy, c.stride)
        g[f'trunk.{2 * li}.bias'] = dy.sum((0, 2, 3))
        if li > 0:
            dy = rb(torch.nn.grad.conv2d_input(acts[li].shape, w[li], dy, c.stride) * (acts[li] > 0))
    return g


@pytest.mark.timeout(240)
@pytest.mark.parametrize('B', [37, 1000])
def test_nature_cnn_tensor_core_gradients_vs_torch(B):
    """Every parameter gradient of the tcgen05 network (data-gradient convolutions, split-K weight gradients,
    masked dense backward): tightly against an fp64 reference that rounds to bf16 where the pipeline does, and
    loosely (direction / norm) against plain fp32 autograd."""
    from xagents_b200.agents import NatureCNN, NatureCnnTc
    torch.manual_seed(2)
    tc = NatureCnnTc(4, 6).cuda()
    with torch.no_grad():
        for p in tc.parameters():
            if p.dim() == 1:
                p.copy_(torch.randn_like(p) * 0.05)
        tc.actor.weight.mul_(30)
    ref = NatureCNN(4, 6).cuda()
    ref.load_state_dict(tc.state_dict())
    x = torch.randint(0, 256, (B, 84, 84, 4), dtype=torch.uint8, device=DEV)
    d_actor = torch.randn((B, 6), device=DEV) / B
    d_critic = torch.randn(B, device=DEV) / B
    actor, critic = tc(x)
    torch.autograd.backward([actor, critic], [d_actor, d_critic])
    ra, rc = ref(x.float() / 255.0)
    torch.autograd.backward([ra, rc.reshape(-1)], [d_actor, d_critic])
    with torch.no_grad():
        emu = _emulated_backward(ref, x, d_actor, d_critic)
    torch.cuda.synchronize()
    report = []
    for (name, p), q in zip(tc.named_parameters(), ref.parameters()):
        assert p.grad.shape == q.grad.shape
        g, r, e = p.grad.flatten().double(), q.grad.flatten().double(), emu[name].flatten()
        report.append((name, float(torch.nn.functional.cosine_similarity(g, r, dim=0)), float((g - r).norm() / r.norm()),
                       float((g - e).norm() / e.norm()), float((g - e).abs().max() / e.abs().max())))
    print('\n'.join(f'{n:18s} vs fp32: cos {c:.5f} rel-L2 {r:.4f} | vs bf16-emulated fp64: rel-L2 {e:.5f} max/max {m:.5f}'
                    for n, c, r, e, m in report))
    for name, cos, rel, rel_emu, max_emu in report:
        # same roundings, different summation order: only bf16 ties that flip (more of them in a larger batch)
        assert rel_emu < 2e-2 and max_emu < 5e-2, f'{name}: {rel_emu:.5f} / {max_emu:.5f} against the bf16-emulated reference'
        # bf16 activations and gradients through five layers against the fp32 network: direction and norm agree
        assert cos > 0.99 and rel < 0.12, f'{name}: cosine {cos:.5f}, relative L2 error {rel:.4f} against fp32'


@pytest.mark.timeout(240)
def test_ppo_fit_with_the_tensor_core_network():
    """PPO.fit end to end with NatureCnnTc: rollout inference, training forward and the whole backward on tcgen05."""
    import importlib.util
    import os

    import numpy as np
    from xagents_b200.agents import PPO, NatureCnnTc, TorchModel
    spec = importlib.util.spec_from_file_location('make_golden', os.path.join(os.path.dirname(__file__), 'golden', 'make_golden.py'))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    T, E, A = 8, 4, 6
    rng = np.random.default_rng(11)
    obs, rewards, dones, resets = mg._streams(rng, 4 * T, E, (84, 84, 4), True, 0.1)
    envs = [mg.ReplayEnv(obs[i], rewards[i], dones[i], resets[i], mg.Discrete(A)) for i in range(E)]
    torch.manual_seed(0)
    net = TorchModel(NatureCnnTc(4, A).cuda())
    before = net.flat_param.clone()
    agent = PPO(envs, net, n_steps=T, mini_batches=4, ppo_epochs=2, quiet=True, seed=5)
    agent.fit(max_steps=3 * T * E)
    torch.cuda.synchronize()
    assert agent.steps == 3 * T * E and net.step == 3 * 2 * 4
    assert torch.isfinite(net.flat_param).all() and not torch.equal(before, net.flat_param)
    losses = torch.stack(agent.loss_history).cpu().numpy()
    assert np.isfinite(losses).all()


@pytest.mark.timeout(120)
def test_graphed_inference_equals_eager_and_follows_weight_updates():
    from xagents_b200.agents import NatureCNN
    from xagents_b200.agents.tc_conv import NatureCnnTcForward
    torch.manual_seed(3)
    net = NatureCNN(4, 6).cuda()
    tc = NatureCnnTcForward(net)
    x = torch.randint(0, 256, (64, 84, 84, 4), dtype=torch.uint8, device=DEV)
    a0, c0 = [t.clone() for t in tc(x)]                  # the outputs are the plan's buffers for this batch size: reused by every call
    run = tc.graphed(64)
    a1, c1 = [t.clone() for t in run(x)]
    torch.cuda.synchronize()
    assert torch.equal(a0, a1) and torch.equal(c0, c1)
    with torch.no_grad():                               # an "optimiser step", then refresh: the graph must see the new weights
        for p in net.parameters():
            p.add_(0.01 * torch.randn_like(p))
    tc.refresh()
    a2, c2 = [t.clone() for t in tc(x)]
    a3, c3 = run(x)
    torch.cuda.synchronize()
    assert torch.equal(a2, a3) and torch.equal(c2, c3) and not torch.equal(a0, a2)


@pytest.mark.timeout(60)
def test_gather_fused_with_scale_and_space_to_depth():
    T, E = 6, 5
    obs = torch.randint(0, 256, (T, E, 84, 84, 4), dtype=torch.uint8, device=DEV)
    idx = torch.randperm(T * E, device=DEV).to(torch.int32)[:17]
    got = ops.gather_s2d_u8_bf16(obs, idx, time_major=(T, E))
    want = ops.space_to_depth_u8_bf16(ops.gather_rows(obs, idx, time_major=(T, E)), 4)
    assert torch.equal(got, want)
    flat = obs.reshape(T * E, 84, 84, 4)
    assert torch.equal(ops.gather_s2d_u8_bf16(flat, idx), ops.space_to_depth_u8_bf16(flat[idx.long()], 4))


@pytest.mark.timeout(120)
@pytest.mark.parametrize('B,H,W,C,kh,kw,N', [(3, 9, 9, 64, 3, 3, 64), (5, 10, 10, 128, 2, 2, 64), (2, 21, 21, 64, 2, 2, 32),
                                             (37, 9, 9, 64, 3, 3, 64), (1, 12, 7, 64, 3, 2, 64), (300, 10, 10, 128, 2, 2, 64)])
def test_conv_weight_gradient_from_natural_nhwc_tensors(B, H, W, C, kh, kw, N):
    """xa_conv_wgrad_nhwc_bf16 (MN-major UMMA operands, dY on the zero-bordered input grid) against fp64 autograd."""
    torch.manual_seed(B + C)
    OH, OW = H - kh + 1, W - kw + 1
    x = torch.randn(B, H, W, C, device=DEV).to(torch.bfloat16)
    dy = torch.randn(B, OH, OW, N, device=DEV).to(torch.bfloat16)
    grid = torch.zeros(B, H, W, N, device=DEV, dtype=torch.bfloat16)
    grid[:, :OH, :OW] = dy
    dw, db = ops.conv_wgrad_nhwc_bf16(x, grid, kh, kw)
    w = torch.zeros(N, C, kh, kw, device=DEV, dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), w)
    y.backward(dy.double().permute(0, 3, 1, 2))
    want = w.grad.permute(0, 2, 3, 1).reshape(N, -1)                       # [N, (kh, kw, c)]
    assert (dw.double() - want).norm() / want.norm() < 2e-5               # bf16 products are exact in fp32; only summation order differs
    want_b = dy.double().sum((0, 1, 2))
    assert (db.double() - want_b).norm() / want_b.norm() < 2e-5


@pytest.mark.timeout(120)
def test_gradients_written_straight_onto_zero_bordered_grids():
    """The backward producers: GEMM column groups, conv output grid, conv unpack -- against the compact result placed by torch."""
    torch.manual_seed(3)
    B = 19
    # FC data gradient [B, 7*7*64] -> 7x7 corner of a [B, 9, 9, 64] grid, with the ReLU mask in compact layout
    dh = torch.randn(B, 512, device=DEV).to(torch.bfloat16)
    wt = (torch.randn(3136, 512, device=DEV) * 0.05).to(torch.bfloat16)
    mask = torch.randn(B, 3136, device=DEV).to(torch.bfloat16)
    compact = ops.gemm_bf16_tn(dh, wt, relu_mask=mask, out_dtype=torch.bfloat16)
    grid = torch.zeros(B, 9, 9, 64, device=DEV, dtype=torch.bfloat16)
    ops.gemm_bf16_tn(dh, wt, relu_mask=mask, out=grid.view(B, -1), col_group=(7 * 64, 9 * 64))
    want = torch.zeros_like(grid)
    want[:, :7, :7] = compact.view(B, 7, 7, 64)
    assert torch.equal(grid, want)
    # conv3 data gradient read from that grid, 9x9 result written on a 10x10 grid
    w3 = (torch.randn(64, 3 * 3 * 64, device=DEV) * 0.05).to(torch.bfloat16)
    m3 = torch.randn(B, 9, 9, 64, device=DEV).to(torch.bfloat16)
    ref = ops.conv2d_nhwc_bf16(compact.view(B, 7, 7, 64), w3, 3, 3, pad=(2, 2), relu_mask=m3)             # [B,9,9,64]
    g2 = torch.zeros(B, 10, 10, 64, device=DEV, dtype=torch.bfloat16)
    ops.conv2d_nhwc_bf16(grid, w3, 3, 3, pad=(2, 2), out_hw=(9, 9), relu_mask=m3, out=g2)
    want2 = torch.zeros_like(g2)
    want2[:, :9, :9] = ref
    assert torch.equal(g2, want2)
    g2f = torch.zeros_like(g2)                                        # the flat kernel (every input pixel fetched once)
    ops.conv2d_nhwc_bf16(grid, w3, 3, 3, pad=(2, 2), out_hw=(9, 9), relu_mask=m3, out=g2f, zero_border=True)
    assert torch.equal(g2f, want2)
    # conv2 data gradient [B,10,10,(dy,dx,32)] unpacked to 20x20 pixels of a 21x21 grid
    w2 = (torch.randn(128, 2 * 2 * 64, device=DEV) * 0.05).to(torch.bfloat16)
    m2 = torch.randn(B, 10, 10, 128, device=DEV).to(torch.bfloat16)
    ref1 = ops.conv2d_nhwc_bf16(ref, w2, 2, 2, pad=(1, 1), relu_mask=m2)                                   # [B,10,10,128]
    g1 = torch.zeros(B, 21, 21, 32, device=DEV, dtype=torch.bfloat16)
    ops.conv2d_nhwc_bf16(g2, w2, 2, 2, pad=(1, 1), out_hw=(10, 10), relu_mask=m2, out=g1, unpack_s2d=True)
    want1 = torch.zeros_like(g1)
    want1[:, :20, :20] = ref1.view(B, 10, 10, 2, 2, 32).permute(0, 1, 3, 2, 4, 5).reshape(B, 20, 20, 32)
    assert torch.equal(g1, want1)
    g1f = torch.zeros_like(g1)
    ops.conv2d_nhwc_bf16(g2, w2, 2, 2, pad=(1, 1), out_hw=(10, 10), relu_mask=m2, out=g1f, unpack_s2d=True, zero_border=True)
    assert torch.equal(g1f, want1)
    # the same two data gradients with the ReLU derivative given as BIT masks (one bit per element of the compact mask tensor)
    g2b, g1b = torch.zeros_like(g2), torch.zeros_like(g1)
    ops.conv2d_nhwc_bf16(grid, w3, 3, 3, pad=(2, 2), out_hw=(9, 9), relu_mask=_pack_bits(m3 > 0), out=g2b, zero_border=True)
    ops.conv2d_nhwc_bf16(g2, w2, 2, 2, pad=(1, 1), out_hw=(10, 10), relu_mask=_pack_bits(m2 > 0), out=g1b, unpack_s2d=True, zero_border=True)
    assert torch.equal(g2b, want2) and torch.equal(g1b, want1)


def _pack_bits(flags):
    """bool tensor -> int32 words, bit j of word i = element 32 i + j (the layout of the kernels' ReLU bit masks)."""
    f = flags.reshape(-1, 32).to(torch.int64)
    words = (f << torch.arange(32, device=f.device, dtype=torch.int64)).sum(1)
    return torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)


@pytest.mark.parametrize('B', [1, 37, 300])
def test_forward_layers_write_their_relu_bit_masks(B):
    """relu_bits_out of the forward convolutions (the first layer from uint8 frames with its 2x2-packed output, and the compact
    64-channel layers): exactly the bits of (output > 0), in the output tensor's own element order."""
    torch.manual_seed(B)
    frames = torch.randint(0, 256, (B, 84, 84, 4), dtype=torch.uint8, device=DEV)
    w1 = (torch.randn(32, 256, device=DEV) * 0.05).bfloat16()
    b1 = torch.randn(32, device=DEV) * 0.1
    bits2 = torch.full((B * 10 * 10 * 128 // 32,), -1, dtype=torch.int32, device=DEV)
    x2 = ops.conv2d_u8_s2d_bf16(frames, w1, 2, 2, bias=b1, relu=True, out_s2d=True, relu_bits_out=bits2)
    assert torch.equal(x2.view(torch.int16), ops.conv2d_u8_s2d_bf16(frames, w1, 2, 2, bias=b1, relu=True, out_s2d=True).view(torch.int16))
    assert torch.equal(bits2, _pack_bits(x2 > 0)) and 0.2 < float((x2 > 0).float().mean()) < 0.8
    x1 = ops.space_to_depth_u8_bf16(frames, 4)
    bits2b = torch.full_like(bits2, -1)
    ops.conv2d_nhwc_bf16(x1, w1, 2, 2, bias=b1, relu=True, out_s2d=True, relu_bits_out=bits2b)
    assert torch.equal(bits2b, bits2)
    w2 = (torch.randn(64, 2 * 2 * 128, device=DEV) * 0.05).bfloat16()
    b2 = torch.randn(64, device=DEV) * 0.1
    bits3 = torch.full((B * 9 * 9 * 64 // 32,), -1, dtype=torch.int32, device=DEV)
    x3 = ops.conv2d_nhwc_bf16(x2, w2, 2, 2, bias=b2, relu=True, relu_bits_out=bits3)
    assert torch.equal(bits3, _pack_bits(x3 > 0)) and 0.2 < float((x3 > 0).float().mean()) < 0.8
    with pytest.raises(Exception, match='ReLU mask bits'):
        ops.conv2d_nhwc_bf16(x2, w2, 2, 2, bias=b2, relu=False, relu_bits_out=bits3)


@pytest.mark.timeout(120)
def test_same_padding_with_zero_border_falls_back_to_the_per_tap_kernel():
    """A 3x3 convolution with pad 1 reads below/right of the image for its last output row/column: the flat kernel cannot
    express that as a row shift, so the zero-border promise must not route it there."""
    torch.manual_seed(5)
    B, H, W, C, N = 6, 9, 9, 64, 64
    x = torch.zeros(B, H, W, C, device=DEV, dtype=torch.bfloat16)
    x[:, :H - 1, :W - 1] = torch.randn(B, H - 1, W - 1, C, device=DEV).to(torch.bfloat16)     # zero last row / column
    w = (torch.randn(N, 3 * 3 * C, device=DEV) * 0.05).to(torch.bfloat16)
    plain = ops.conv2d_nhwc_bf16(x, w, 3, 3, pad=(1, 1))
    promised = ops.conv2d_nhwc_bf16(x, w, 3, 3, pad=(1, 1), zero_border=True)
    assert torch.equal(plain, promised)
    want = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().reshape(N, 3, 3, C).permute(0, 3, 1, 2), padding=1)
    assert float((plain.float() - want.permute(0, 2, 3, 1)).abs().max()) <= 4e-3 * float(want.abs().max())
