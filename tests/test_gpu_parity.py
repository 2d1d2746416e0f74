"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the reference-made goldens.

Run on the B200 box: python -m pytest tests -m gpu.  Tolerances (BASELINE.json north_star):
bit-exact for gathers / index work; <= 1e-5 relative (magnitude-normalised) for
returns, advantages, losses and gradients.
"""
import numpy as np
import pytest
import torch

import oracle
from xagents_b200 import ops, synthetic

pytestmark = pytest.mark.gpu

REL = 1e-5
DEV = 'cuda:0'


def cu(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


def close(got, want, scale=None, rel=REL):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    want = np.asarray(want)
    s = float(np.max(np.abs(want))) if scale is None else float(scale)
    err = float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64)))) if want.size else 0.0
    assert err <= rel * max(s, 1e-30), f'max abs err {err:.3e} > {rel} * {s:.3e}'


# ---------------------------------------------------------------------------------------------- returns
@pytest.mark.parametrize('case', ['kat_returns', 'ppo_image', 'ppo_cartpole', 'ppo_ragged', 'ppo_single_env', 'ppo_box', 'ppo_softmax'])
def test_gae_sequential_bit_exact_vs_reference_golden(golden, case):
    g = golden(case)
    if case == 'kat_returns':
        T, E = g['rewards'].shape
        want = g['ppo_returns']
    else:
        T, E = int(g['n_steps']), int(g['n_envs'])
        want = g['returns'].reshape(T, E)
    got = ops.gae_returns(cu(g['rewards'].reshape(T, E)), cu(g['values'].reshape(T, E)), cu(g['next_values'].reshape(E)),
                          cu(g['dones'].reshape(T + 1, E)), float(g['gamma']), float(g['lam']), mode='sequential')
    assert np.array_equal(got.cpu().numpy(), want)
    for mode in ('auto', 'chunked'):
        got = ops.gae_returns(cu(g['rewards'].reshape(T, E)), cu(g['values'].reshape(T, E)),
                              cu(g['next_values'].reshape(E)), cu(g['dones'].reshape(T + 1, E)), float(g['gamma']),
                              float(g['lam']), mode=mode)
        close(got, want)


@pytest.mark.parametrize('case', ['kat_returns', 'a2c_image', 'a2c_vector', 'a2c_box', 'a2c_softmax'])
def test_nstep_sequential_bit_exact_vs_reference_golden(golden, case):
    g = golden(case)
    want = g['a2c_returns'] if case == 'kat_returns' else g['returns']
    T, E = want.shape
    for mode in ('sequential', 'auto', 'chunked'):
        got = ops.nstep_returns(cu(g['rewards']), cu(g['dones']), cu(g['next_values']), float(g['gamma']), mode=mode)
        if mode == 'sequential':
            assert np.array_equal(got.cpu().numpy(), want)
        else:
            close(got, want)


@pytest.mark.parametrize('T,E,p_done,gamma,lam', [
    (128, 256, 0.01, 0.99, 0.95), (5, 16, 0.1, 0.99, 0.95), (2048, 16, 0.01, 0.99, 0.95), (2048, 64, 0.0, 0.999, 0.99),
    (333, 37, 0.5, 0.9, 0.8), (1, 1, 0.0, 0.99, 0.95), (17, 1, 0.2, 0.99, 0.95), (64, 4099, 0.02, 1.0, 1.0),
    (9, 65536, 0.01, 0.99, 0.95),
])
def test_gae_all_modes_vs_oracle(T, E, p_done, gamma, lam):
    ro = synthetic.make_rollout(T, E, with_obs=False, p_done=p_done, epochs=0, seed=T * 7919 + E)
    want, want_adv = oracle.gae_returns(ro.rewards, ro.dones, ro.values, ro.last_values, gamma, lam, return_advantages=True)
    truth = oracle.gae_returns(ro.rewards, ro.dones, ro.values, ro.last_values, gamma, lam, dtype=np.float64)
    args = (cu(ro.rewards), cu(ro.values), cu(ro.last_values), cu(ro.dones), gamma, lam)
    got, adv = ops.gae_returns(*args, mode='sequential', with_advantages=True)
    assert np.array_equal(got.cpu().numpy(), want) and np.array_equal(adv.cpu().numpy(), want_adv)
    for mode in ('auto', 'chunked'):
        got, adv = ops.gae_returns(*args, mode=mode, with_advantages=True)
        close(got, want)
        close(got, truth)
        close(adv, want_adv, scale=np.abs(want).max())
    want = oracle.nstep_returns(ro.rewards, ro.dones, ro.last_values, gamma)
    nargs = (cu(ro.rewards), cu(ro.dones), cu(ro.last_values), gamma)
    assert np.array_equal(ops.nstep_returns(*nargs, mode='sequential').cpu().numpy(), want)
    for mode in ('auto', 'chunked'):
        close(ops.nstep_returns(*nargs, mode=mode), want)


def test_gae_properties():
    T, E = 96, 200
    ro = synthetic.make_rollout(T, E, with_obs=False, epochs=0)
    # every step terminal -> returns == rewards (advantage = r - V, + V)
    dones = np.ones_like(ro.dones)
    got = ops.gae_returns(cu(ro.rewards), cu(ro.values), cu(ro.last_values), cu(dones), 0.99, 0.95, mode='chunked')
    close(got, ro.rewards, rel=1e-6)
    # gamma = lam = 1, no dones -> suffix sums of rewards + bootstrap
    dones = np.zeros_like(ro.dones)
    got = ops.gae_returns(cu(ro.rewards), cu(ro.values), cu(ro.last_values), cu(dones), 1.0, 1.0, mode='chunked')
    want = np.cumsum(ro.rewards[::-1].astype(np.float64), axis=0)[::-1] + ro.last_values
    close(got, want)
    got = ops.nstep_returns(cu(ro.rewards), cu(dones), cu(ro.last_values), 1.0, mode='chunked')
    close(got, want)


# ---------------------------------------------------------------------------------------------- gathers
@pytest.mark.parametrize('mode', ['bulk', 'vector', 'auto'])
@pytest.mark.parametrize('T,E,shape,dtype', [
    (16, 8, (84, 84, 4), np.uint8), (5, 16, (84, 84, 1), np.uint8), (128, 16, (4,), np.float32), (7, 3, (5,), np.float32),
    (9, 5, (33,), np.uint8), (4, 4, (70000,), np.uint8), (3, 2, (16,), np.uint8),
])
def test_gather_rows_bit_exact(mode, T, E, shape, dtype):
    rng = np.random.default_rng(T * 131 + E)
    n = T * E
    row_bytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    if mode == 'bulk' and row_bytes % 16:
        pytest.skip('bulk path needs 16-byte rows')
    obs = (rng.integers(0, 256, size=(T, E) + shape).astype(dtype) if dtype == np.uint8
           else rng.standard_normal((T, E) + shape).astype(dtype))
    flat = oracle.concat_step_batches(obs)[0]
    perm = rng.permutation(n).astype(np.int32)
    for idx in (perm, perm[: n // 3 + 1], rng.integers(0, n, size=2 * n + 5).astype(np.int32)):
        want = np.take(flat, idx, axis=0)
        got = ops.gather_rows(cu(obs), cu(idx), time_major=(T, E), mode=mode)
        assert got.dtype == torch.as_tensor(obs).dtype and np.array_equal(got.cpu().numpy(), want)
        got = ops.gather_rows(cu(flat), cu(idx), mode=mode)          # already-flat source, no remap
        assert np.array_equal(got.cpu().numpy(), want)


def test_gather_empty_and_errors():
    from xagents_b200._ffi import XAError
    obs = torch.zeros((4, 2, 16), dtype=torch.uint8, device=DEV)
    out = ops.gather_rows(obs, torch.zeros(0, dtype=torch.int32, device=DEV), time_major=(4, 2))
    assert out.shape == (0, 16)
    with pytest.raises(XAError):                                     # misaligned rows cannot take the bulk path
        ops.gather_rows(torch.zeros((4, 2, 17), dtype=torch.uint8, device=DEV),
                        torch.zeros(3, dtype=torch.int32, device=DEV), time_major=(4, 2), mode='bulk')
    with pytest.raises(TypeError):
        ops.gather_rows(obs, torch.zeros(3, dtype=torch.int64, device=DEV))
    with pytest.raises(ValueError):                                  # host tensors are refused: no CPU path
        ops.gather_rows(obs.cpu(), torch.zeros(3, dtype=torch.int32))


@pytest.mark.parametrize('case', ['ppo_image', 'ppo_cartpole', 'ppo_ragged', 'ppo_single_env'])
def test_minibatch_gather_vs_reference_golden(golden, case):
    """Every minibatch the reference's get_mini_batches produced, from the time-major rollout."""
    g = golden(case)
    T, E, B = int(g['n_steps']), int(g['n_envs']), int(g['mini_batch_size'])
    image = g['flat_states'].ndim > 2
    obs_tm = g['states_time_major'].astype(np.uint8) if image else g['states_time_major']
    tm = lambda flat: np.ascontiguousarray(flat.reshape(E, T).T)
    fields = [tm(g['flat_actions']), tm(g['flat_returns']), tm(g['flat_values']), tm(g['flat_log_probs'])]
    obs_d, fields_d = cu(obs_tm), [cu(f) for f in fields]
    k = 0
    for perm in g['shuffles']:
        for lo, hi in oracle.minibatch_slices(T * E, B):
            idx = cu(perm[lo:hi].astype(np.int32))
            got_obs, got_fields = ops.gather_minibatch(obs_d, fields_d, idx, time_major=(T, E))
            want = g[f'mb{k}_states']
            assert np.array_equal(got_obs.cpu().numpy().astype(np.float32), want)
            for got, name in zip(got_fields, ('actions', 'returns', 'old_values', 'old_log_probs')):
                assert np.array_equal(got.cpu().numpy(), g[f'mb{k}_{name}'].reshape(-1)), name
            if image:
                scaled = ops.gather_rows_scaled(obs_d, idx, time_major=(T, E))
                assert np.array_equal(scaled.cpu().numpy(), oracle.scale_images(want))
            k += 1
    assert k == len(g['losses'])


def test_gather_full_size_is_a_permutation():
    """C3-sized epoch (N = 32768 rows of 28224 B): permutation properties instead of a CPU copy."""
    T, E = 128, 256
    N = T * E
    obs = torch.randint(0, 256, (T, E, 84, 84, 4), dtype=torch.uint8, device=DEV)
    perm = torch.randperm(N, device=DEV).to(torch.int32)
    for mode in ('bulk', 'vector'):
        out = ops.gather_rows(obs, perm, time_major=(T, E), mode=mode)
        # checksum of checksums: per-row sums must be the permuted per-row sums
        rows = obs.view(T * E, -1)
        src_rows = (perm.long() % T) * E + perm.long() // T
        assert torch.equal(out.view(N, -1).sum(1, dtype=torch.int64), rows.sum(1, dtype=torch.int64)[src_rows])
        # inverse permutation round trip restores the env-major flat tensor exactly
        inv = torch.empty_like(perm)
        inv[perm.long()] = torch.arange(N, device=DEV, dtype=torch.int32)
        back = ops.gather_rows(out, inv, mode=mode)
        assert torch.equal(back.view(E, T, -1).transpose(0, 1).reshape(T, E, 84, 84, 4), obs)
        del out, back


# ---------------------------------------------------------------------------------------------- losses
def _loss_inputs(g, k):
    return (g[f'mb{k}_actor'], g[f'mb{k}_critic'].reshape(-1), g[f'mb{k}_actions'].reshape(-1),
            g[f'mb{k}_old_log_probs'].reshape(-1), g[f'mb{k}_old_values'].reshape(-1), g[f'mb{k}_returns'].reshape(-1))


@pytest.mark.parametrize('case', ['ppo_image', 'ppo_cartpole', 'ppo_ragged', 'ppo_single_env'])
def test_ppo_loss_vs_reference_golden(golden, case):
    g = golden(case)
    hp = dict(clip_norm=float(g['clip_norm']), entropy_coef=float(g['entropy_coef']),
              value_loss_coef=float(g['value_loss_coef']), advantage_epsilon=float(g['advantage_epsilon']))
    for k, ref_loss in enumerate(g['losses']):
        actor, critic, actions, old_lp, old_v, ret = _loss_inputs(g, k)
        n = len(ret)
        mom = ops.adv_moments(cu(ret), cu(old_v), None, [0, n])
        sc, d_actor, d_values, adv = ops.ppo_loss(cu(actor), cu(critic), cu(actions), cu(old_lp), cu(old_v), cu(ret),
                                                 moments=mom[0], return_advantages=True, **hp)
        sc = sc.cpu().numpy()
        scale = abs(float(ref_loss))
        assert abs(sc[0] - ref_loss) <= REL * scale, (k, sc, ref_loss)
        _, ref_ent, ref_vl, ref_pg = g['means'][k]
        assert abs(sc[1] - ref_pg) <= REL * scale and abs(sc[2] - 0.5 * ref_vl) <= REL * scale
        assert abs(sc[3] - ref_ent) <= REL * scale
        ref_adv = g[f'mb{k}_advantages'].reshape(-1)
        close(adv, ref_adv)
        # gradients: closed form of the oracle and torch autograd (stand-in for the tape)
        dl, dv = oracle.ppo_loss_grads(actor, critic, actions, old_v, ret, old_lp, ref_adv, hp['clip_norm'],
                                       hp['entropy_coef'], hp['value_loss_coef'])
        close(d_actor, dl)
        close(d_values, dv)
        # supplying the reference's own normalised advantages must give the same thing
        sc2, da2, dv2, _ = ops.ppo_loss(cu(actor), cu(critic), cu(actions), cu(old_lp), cu(old_v), cu(ret),
                                        advantages=cu(ref_adv), **hp)
        assert abs(sc2.cpu().numpy()[0] - ref_loss) <= REL * scale
        close(da2, dl)


def _normal_grads_fp64(actor, critic, actions, old_lp, old_v, ret, adv, clip, ec, vc, ppo=True):
    """d loss / d(loc, V) for the MultivariateNormalDiag branch by fp64 autograd with TF's tie rules (oracle/torch_ref.py
    conventions); the closed forms of oracle/hotpath.py cover the two Categorical branches."""
    from oracle.torch_ref import _tf_clip, _tf_max
    t = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64)
    loc, v = t(actor).clone().requires_grad_(True), t(critic).clone().requires_grad_(True)
    a, old_v, ret = t(actions).reshape(loc.shape), t(old_v), t(ret)
    k = loc.shape[-1]
    logp = -0.5 * ((a - loc) ** 2).sum(-1) - 0.5 * k * np.log(2 * np.pi)
    entropy = torch.full_like(logp, 0.5 * k * (1 + np.log(2 * np.pi))).mean()
    if ppo:
        old_lp, adv = t(old_lp), t(adv)
        clipped = old_v + _tf_clip(v - old_v, -clip, clip)
        vl = 0.5 * _tf_max((v - ret) ** 2, (clipped - ret) ** 2).mean()
        ratio = torch.exp(logp - old_lp)
        pg = _tf_max(-adv * ratio, -adv * _tf_clip(ratio, 1 - clip, 1 + clip)).mean()
    else:
        vl = ((v - ret) ** 2).mean()
        pg = -((ret - old_v) * logp).mean()
    (pg - entropy * ec + vl * vc).backward()
    return loc.grad.numpy(), v.grad.numpy()


@pytest.mark.parametrize('case', ['ppo_box', 'ppo_softmax'])
def test_ppo_loss_other_distribution_branches_vs_reference_golden(golden, case):
    """MultivariateNormalDiag (Box actions [N, k]) and Categorical(probs=) (softmax-output model): the other two
    branches of A2C.get_distribution (a2c/agent.py:50-63), fixtures from the reference's own PPO.train_step."""
    g = golden(case)
    kind = str(g['actor_kind'])
    assert kind == {'ppo_box': 'normal', 'ppo_softmax': 'probs'}[case]
    hp = dict(clip_norm=float(g['clip_norm']), entropy_coef=float(g['entropy_coef']),
              value_loss_coef=float(g['value_loss_coef']), advantage_epsilon=float(g['advantage_epsilon']))
    for k, ref_loss in enumerate(g['losses']):
        actor, critic = g[f'mb{k}_actor'], g[f'mb{k}_critic'].reshape(-1)
        actions = g[f'mb{k}_actions'].reshape(actor.shape) if kind == 'normal' else g[f'mb{k}_actions'].reshape(-1)
        old_lp, old_v, ret = (g[f'mb{k}_{name}'].reshape(-1) for name in ('old_log_probs', 'old_values', 'returns'))
        mom = ops.adv_moments(cu(ret), cu(old_v), None, [0, len(ret)])
        sc, d_actor, d_values, adv = ops.ppo_loss(cu(actor), cu(critic), cu(actions), cu(old_lp), cu(old_v), cu(ret),
                                                 moments=mom[0], actor_kind=kind, return_advantages=True, **hp)
        sc = sc.cpu().numpy()
        _, ref_ent, ref_vl, ref_pg = g['means'][k]
        scale = max(abs(float(ref_loss)), abs(float(ref_ent)) * hp['entropy_coef'])
        assert abs(sc[0] - ref_loss) <= REL * scale, (k, sc, ref_loss)
        assert abs(sc[1] - ref_pg) <= REL * scale and abs(sc[2] - 0.5 * ref_vl) <= REL * scale
        assert abs(sc[3] - ref_ent) <= REL * max(scale, abs(float(ref_ent)))
        ref_adv = g[f'mb{k}_advantages'].reshape(-1)
        close(adv, ref_adv)
        if kind == 'probs':
            dl, dv = oracle.ppo_loss_grads(actor, critic, actions, old_v, ret, old_lp, ref_adv, hp['clip_norm'],
                                           hp['entropy_coef'], hp['value_loss_coef'], is_probs=True)
        else:
            dl, dv = _normal_grads_fp64(actor, critic, actions, old_lp, old_v, ret, ref_adv, hp['clip_norm'],
                                        hp['entropy_coef'], hp['value_loss_coef'])
        close(d_actor, dl)
        close(d_values, dv)


@pytest.mark.parametrize('case', ['a2c_box', 'a2c_softmax'])
def test_a2c_loss_other_distribution_branches_vs_reference_golden(golden, case):
    g = golden(case)
    kind = str(g['actor_kind'])
    actor, critic = g['actor'], g['critic'].reshape(-1)
    actions = g['flat_actions'].reshape(actor.shape) if kind == 'normal' else g['flat_actions'].reshape(-1)
    old_v, ret = g['flat_values'].reshape(-1), g['flat_returns'].reshape(-1)
    ec, vc = float(g['entropy_coef']), float(g['value_loss_coef'])
    sc, d_actor, d_values = ops.a2c_loss(cu(actor), cu(critic), cu(actions), cu(old_v), cu(ret), entropy_coef=ec,
                                         value_loss_coef=vc, actor_kind=kind)
    ref = float(g['loss'][0])
    assert abs(sc.cpu().numpy()[0] - ref) <= REL * abs(ref)
    if kind == 'probs':
        dl, dv = oracle.a2c_loss_grads(actor, critic, actions, old_v, ret, ec, vc, is_probs=True)
    else:
        dl, dv = _normal_grads_fp64(actor, critic, actions, None, old_v, ret, None, 0.0, ec, vc, ppo=False)
    close(d_actor, dl)
    close(d_values, dv)


@pytest.mark.parametrize('case', ['a2c_image', 'a2c_vector'])
def test_a2c_loss_vs_reference_golden(golden, case):
    g = golden(case)
    actor, critic = g['actor'], g['critic'].reshape(-1)
    actions, old_v, ret = g['flat_actions'].reshape(-1), g['flat_values'].reshape(-1), g['flat_returns'].reshape(-1)
    ec, vc = float(g['entropy_coef']), float(g['value_loss_coef'])
    sc, d_actor, d_values = ops.a2c_loss(cu(actor), cu(critic), cu(actions), cu(old_v), cu(ret), entropy_coef=ec,
                                         value_loss_coef=vc)
    ref = float(g['loss'][0])
    assert abs(sc.cpu().numpy()[0] - ref) <= REL * abs(ref)
    dl, dv = oracle.a2c_loss_grads(actor, critic, actions, old_v, ret, ec, vc)
    close(d_actor, dl)
    close(d_values, dv)


def test_kat3_kat4_on_device(golden):
    g = golden('kat_returns')
    idx = np.array([7, 2, 9, 0, 5, 11], np.int32)
    logits = np.float32([[.1, -.2, .3], [1, 0, -1], [.5, .5, .5], [-.3, .8, .2], [2, -1, 0], [0, .1, -.1]])
    new_v = np.float32([.35, .6, -.45, .55, .25, .1])
    actions = np.float32([0, 2, 1, 1, 0, 2])
    old_logp = np.float32([-1, -2.3, -1.2, -.7, -.3, -1.05])
    # scalars are read THROUGH the indices from the time-major [T,E] buffers (fused gather)
    T, E = 4, 3

    def full(mb):                      # scatter the minibatch values into a time-major [T,E] field
        flat = np.zeros(T * E, np.float32)
        flat[idx] = mb
        return np.ascontiguousarray(flat.reshape(E, T).T)

    ret_tm, val_tm = g['ppo_returns'], g['values']
    mom = ops.adv_moments(cu(ret_tm), cu(val_tm), cu(idx), [0, 6], time_major=(T, E))
    sc, d_actor, d_values, adv = ops.ppo_loss(cu(logits), cu(new_v), cu(full(actions)), cu(full(old_logp)), cu(val_tm),
                                             cu(ret_tm), idx=cu(idx), time_major=(T, E), moments=mom[0],
                                             return_advantages=True)
    np.testing.assert_allclose(adv.cpu().numpy(), [-0.09828994, 0.45400777, -0.5042908, 0.6545685, 1.3239344, -1.8299301], atol=2e-6)
    np.testing.assert_allclose(sc.cpu().numpy(), [0.521738112, -0.044882495, 1.152010560, 0.938469231], atol=2e-6)
    np.testing.assert_allclose(d_values.cpu().numpy(), [0.045666665, 0, 0, -0.11164874, -0.20891115, 0.175], atol=1e-6)
    np.testing.assert_allclose(d_actor.cpu().numpy()[0], [9.9536581e-03, -3.8876350e-03, -6.0660210e-03], atol=1e-7)
    sc, _, _ = ops.a2c_loss(cu(logits), cu(new_v), cu(full(actions)), cu(val_tm), cu(ret_tm), idx=cu(idx), time_major=(T, E))
    np.testing.assert_allclose(sc.cpu().numpy()[:3], [1.288439870, 0.216375038, 2.162899017], atol=2e-6)


@pytest.mark.parametrize('kind,A', [('logits', 6), ('probs', 6), ('logits', 18), ('probs', 3), ('normal', 4), ('logits', 2)])
def test_loss_kinds_vs_autograd(kind, A):
    from oracle import torch_ref
    rng = np.random.default_rng(A * 17 + len(kind))
    n = 5000
    logits = rng.standard_normal((n, A)).astype(np.float32)
    old_v = rng.standard_normal(n).astype(np.float32)
    new_v = (old_v + 0.15 * rng.standard_normal(n)).astype(np.float32)
    ret = rng.standard_normal(n).astype(np.float32)
    adv = oracle.normalize_advantages(ret, old_v, 1e-8)
    if kind == 'normal':
        actions = (logits + rng.standard_normal((n, A))).astype(np.float32)
        logp, _ = oracle.diag_normal_logp_entropy(logits, actions)
        actor = logits
    else:
        actions = rng.integers(0, A, n).astype(np.float32)
        actor = np.exp(logits - logits.max(-1, keepdims=True))
        actor = (actor / actor.sum(-1, keepdims=True)).astype(np.float32) if kind == 'probs' else logits
        logp, _, _ = oracle.categorical_logp_entropy(actor, actions, kind == 'probs')
    old_lp = (logp + 0.15 * rng.standard_normal(n)).astype(np.float32)
    mom = ops.adv_moments(cu(ret), cu(old_v), None, [0, n])
    sc, d_actor, d_values, adv_d = ops.ppo_loss(cu(actor), cu(new_v), cu(actions), cu(old_lp), cu(old_v), cu(ret),
                                               moments=mom[0], actor_kind=kind, return_advantages=True)
    close(adv_d, adv)
    if kind == 'normal':
        a_t, lp_t = torch.tensor(actor, requires_grad=True), None
        v_t = torch.tensor(new_v, requires_grad=True)
        lp = -0.5 * ((torch.tensor(actions) - a_t) ** 2).sum(-1) - 0.5 * A * np.log(2 * np.pi)
        ent = torch.full((n,), 0.5 * A * (1 + np.log(2 * np.pi)))
        ratio = torch.exp(lp - torch.tensor(old_lp))
        advt = torch.tensor(adv)
        s1, s2 = -advt * ratio, -advt * torch_ref._tf_clip(ratio, 0.9, 1.1)
        vc = torch.tensor(old_v) + torch_ref._tf_clip(v_t - torch.tensor(old_v), -0.1, 0.1)
        rt = torch.tensor(ret)
        loss = (torch_ref._tf_max(s1, s2).mean() - 0.01 * ent.mean()
                + 0.5 * 0.5 * torch_ref._tf_max((v_t - rt) ** 2, (vc - rt) ** 2).mean())
        loss.backward()
        want_sc, da, dv = dict(loss=loss.item()), a_t.grad.numpy(), v_t.grad.numpy()
    else:
        want_sc, da, dv = torch_ref.ppo_loss_autograd(actor, new_v, actions, old_v, ret, old_lp, adv, .1, .01, .5, kind == 'probs')
    assert abs(sc.cpu().numpy()[0] - want_sc['loss']) <= REL * abs(want_sc['loss'])
    close(d_actor, da)
    close(d_values, dv)
    if kind != 'normal':
        want_sc, da, dv = torch_ref.a2c_loss_autograd(actor, new_v, actions, old_v, ret, .01, .5, kind == 'probs')
        sc, d_actor, d_values = ops.a2c_loss(cu(actor), cu(new_v), cu(actions), cu(old_v), cu(ret), actor_kind=kind)
        assert abs(sc.cpu().numpy()[0] - want_sc['loss']) <= REL * abs(want_sc['loss'])
        close(d_actor, da)
        close(d_values, dv)


def test_moment_parts_combine_like_one_batch():
    """Sharded moments (what ranks all-gather) combine to the whole-minibatch normalisation."""
    rng = np.random.default_rng(3)
    n, A = 8192, 6
    ret = (3 + 2 * rng.standard_normal(n)).astype(np.float32)
    old_v = rng.standard_normal(n).astype(np.float32)
    logits = rng.standard_normal((n, A)).astype(np.float32)
    actions = rng.integers(0, A, n).astype(np.float32)
    new_v, old_lp = old_v.copy(), np.full(n, -1.5, np.float32)
    whole = ops.adv_moments(cu(ret), cu(old_v), None, [0, n])
    parts = ops.adv_moments(cu(ret), cu(old_v), None, [0, 1000, 1000, 5000, n])       # incl. an empty part
    args = (cu(logits), cu(new_v), cu(actions), cu(old_lp), cu(old_v), cu(ret))
    _, _, _, adv1 = ops.ppo_loss(*args, moments=whole[0], return_advantages=True)
    _, _, _, adv2 = ops.ppo_loss(*args, moments=parts, return_advantages=True)
    _, _, _, adv3 = ops.ppo_loss(*args, moments=(parts, 0, 4, 4), return_advantages=True)
    close(adv2, adv1.cpu().numpy(), rel=1e-6)
    assert torch.equal(adv2, adv3)
    close(adv1, oracle.normalize_advantages(ret, old_v, 1e-8))
    close(adv1, oracle.normalize_advantages(ret, old_v, 1e-8, dtype=np.float64))


def test_loss_is_deterministic_and_workspace_reusable():
    ro = synthetic.make_rollout(128, 64, with_obs=False, epochs=1)
    N = ro.batch_size
    ret = oracle.gae_returns(ro.rewards, ro.dones, ro.values, ro.last_values, 0.99, 0.95)
    flat = lambda x: oracle.concat_step_batches(x)[0].reshape(-1)
    args = (cu(ro.new_logits), cu(ro.new_values), cu(flat(ro.actions)), cu(flat(ro.log_probs)), cu(flat(ro.values)), cu(flat(ret)))
    ws = ops.loss_workspace(N, DEV)
    mom = ops.adv_moments(args[5], args[4], None, [0, N])
    first = None
    for _ in range(5):
        sc, da, dv, _ = ops.ppo_loss(*args, moments=mom[0], workspace=ws)
        cur = (sc.cpu().numpy().copy(), da.cpu().numpy().copy())
        if first is None:
            first = cur
        assert np.array_equal(cur[0], first[0]) and np.array_equal(cur[1], first[1])
    assert int(ws.view(torch.int32)[0]) == 0          # ticket restored


# ---------------------------------------------------------------------------------------------- whole path
@pytest.mark.parametrize('T,E,mb', [(128, 16, 4), (16, 8, 4), (7, 3, 4)])
def test_ppo_train_step_vs_oracle(T, E, mb):
    """GAE -> per-minibatch (gather, moments, loss) driven from time-major buffers vs the oracle's
    restatement of PPO.train_step with identical permutations."""
    ro = synthetic.make_rollout(T, E, obs_shape=(12, 12, 4), epochs=2, p_done=0.05)
    want = oracle.ppo_train_step(ro.obs, ro.rewards, ro.dones, ro.values, ro.last_values, ro.actions, ro.log_probs,
                                 ro.permutations, ro.new_logits, ro.new_values, mini_batches=mb, keep_states=True)
    d = {k: cu(getattr(ro, k)) for k in ('obs', 'rewards', 'values', 'last_values', 'dones', 'actions', 'log_probs',
                                         'new_logits', 'new_values')}
    ret = ops.gae_returns(d['rewards'], d['values'], d['last_values'], d['dones'], 0.99, 0.95, mode='sequential')
    assert np.array_equal(ret.cpu().numpy(), want['returns'])
    N, B = T * E, (T * E) // mb
    k = 0
    for perm in ro.permutations:
        perm_d = cu(perm)
        slices = oracle.minibatch_slices(N, B)
        mom = ops.adv_moments(ret, d['values'], perm_d, [s[0] for s in slices] + [N], time_major=(T, E))
        for m, (lo, hi) in enumerate(slices):
            idx = perm_d[lo:hi]
            w = want['minibatches'][k]
            states = ops.gather_rows(d['obs'], idx, time_major=(T, E))
            assert np.array_equal(states.cpu().numpy(), w['states'])
            logits_mb = ops.gather_rows(d['new_logits'], idx)
            values_mb = ops.gather_rows(d['new_values'].view(-1, 1), idx).view(-1)
            sc, da, dv, adv = ops.ppo_loss(logits_mb, values_mb, d['actions'], d['log_probs'], d['values'], ret, idx=idx,
                                           time_major=(T, E), moments=mom[m], return_advantages=True)
            sc = sc.cpu().numpy()
            # the total can cancel to ~0 while its terms are O(1): normalise by the largest term
            scale = max(abs(float(w[name])) for name in ('loss', 'pg', 'vl', 'entropy'))
            for j, name in enumerate(('loss', 'pg', 'vl', 'entropy')):
                assert abs(sc[j] - w[name]) <= REL * scale, (k, name, sc[j], w[name])
            close(adv, w['advantages'])
            close(da, w['dlogits'])
            close(dv, w['dvalues'])
            k += 1
    assert k == len(want['minibatches'])


# ---------------------------------------------------------------------------------------------- optimiser
@pytest.mark.parametrize('n,clip', [(1687719, 0.5), (1000, None), (4099, 10.0)])
def test_clip_adam_vs_oracle(n, clip):
    rng = np.random.default_rng(n)
    p = rng.standard_normal(n).astype(np.float32)
    m = np.zeros(n, np.float32)
    v = np.zeros(n, np.float32)
    pd, md, vd = cu(p), cu(m), cu(v)
    ws = ops.optim_workspace(DEV)
    for step in (1, 2, 3):
        g = (rng.standard_normal(n) * 0.01).astype(np.float32)
        gg = g
        if clip:
            (gg,), norm = oracle.clip_by_global_norm([g], clip)
        p, m, v = oracle.adam_step(p, gg, m, v, step)
        ops.clip_adam(pd, cu(g), md, vd, step, workspace=ws, clip_norm=clip)
        close(pd, p, rel=2e-6)
        close(md, m, rel=1e-5)
        close(vd, v, rel=1e-5)


# ---------------------------------------------------------------------------------------------- PPOHotPath
@pytest.mark.parametrize('fuse,chunk,overlap,staging', [(True, None, True, 2), (False, None, True, 2), (True, 1, True, 2),
                                                        (False, 1, False, 1), (False, 8, True, 1), (True, 3, True, 2),
                                                        (False, 'schedule', True, 2)])
@pytest.mark.parametrize('T,E,mb', [(16, 8, 4), (7, 3, 4)])
def test_hotpath_pipeline_vs_oracle(T, E, mb, fuse, chunk, overlap, staging):
    """The prepared two-stream pipeline (what bench.py times) against the oracle's PPO.train_step,
    including the trailing short minibatch (7*3 % 4 != 0) and every staging / granularity layout."""
    _check_pipeline(T, E, mb, fuse, chunk, overlap, staging, 'event')


@pytest.mark.parametrize('sync', ['progress', 'progress-memop'])
@pytest.mark.parametrize('chunk,staging', [(None, 2), (3, 2), (1, 1), ('schedule', 2), (5, 3)])
@pytest.mark.parametrize('T,E,mb', [(16, 8, 4), (7, 3, 4)])
def test_hotpath_progress_sync_vs_oracle(T, E, mb, chunk, staging, sync):
    """The same pipeline with per-minibatch completion counters (the gather kernel publishes finished rows; the compute stream
    waits on its minibatch's counter with a one-warp spin kernel, or with cuStreamWaitValue32) instead of one event per launch."""
    _check_pipeline(T, E, mb, True, chunk, True, staging, sync)


def _check_pipeline(T, E, mb, fuse, chunk, overlap, staging, sync):
    from xagents_b200.hotpath import PPOHotPath
    K = 2
    ro = synthetic.make_rollout(T, E, obs_shape=(84, 84, 4), epochs=K, p_done=0.05, seed=T + E)
    want = oracle.ppo_train_step(ro.obs, ro.rewards, ro.dones, ro.values, ro.last_values, ro.actions, ro.log_probs,
                                 ro.permutations, ro.new_logits, ro.new_values, mini_batches=mb, keep_states=True)
    if chunk == 'schedule':                                           # uneven explicit launch schedule
        n_mb = K * -(-(T * E) // ((T * E) // mb))
        chunk = [3, 1, 2] + [n_mb - 6]
    hp = PPOHotPath(T, E, (84, 84, 4), ro.n_actions, ppo_epochs=K, mini_batches=mb, device=DEV, fuse_fields=fuse,
                    gather_chunk=chunk, overlap=overlap, staging=staging, scan_mode='sequential', sync=sync)
    assert hp.sync == sync
    hp.load(ro)
    hp.perms.copy_(torch.as_tensor(np.stack(ro.permutations)))
    for i, w in enumerate(want['minibatches']):                       # the "forward pass" outputs per minibatch
        n = len(w['idx'])
        hp.actor_out[i, :n].copy_(torch.as_tensor(ro.new_logits[w['idx']]))
        hp.critic_out[i, :n].copy_(torch.as_tensor(ro.new_values[w['idx']]))
    hp.prepare()
    seen = {}

    def after_loss(i):                                                # runs in stream order right after loss i
        n = hp.mb_rows[i]
        states = hp.minibatch_views(i)[0]
        seen[i] = (hp.d_actor[:n].clone(), hp.d_values[:n].clone(), states.clone(),
                   [f.clone() for f in hp.minibatch_views(i)[1:]])

    for _ in range(3):                                                # repeatedly: buffers, tickets and counters must be reusable
        seen.clear()
        hp.run(after_loss=after_loss)
        torch.cuda.synchronize()
        assert getattr(hp, 'wait_status', None) is None or int(hp.wait_status.item()) == 0      # no device-side wait gave up
        assert np.array_equal(hp.returns.cpu().numpy(), want['returns'])
        sc = hp.scalars.cpu().numpy()
        assert len(want['minibatches']) == hp.n_mb
        for i, w in enumerate(want['minibatches']):
            scale = max(abs(float(w[name])) for name in ('loss', 'pg', 'vl', 'entropy'))
            for j, name in enumerate(('loss', 'pg', 'vl', 'entropy')):
                assert abs(sc[i, j] - w[name]) <= REL * scale, (i, name, sc[i, j], w[name])
            da, dv, states, fields = seen[i]
            close(da, w['dlogits'])
            close(dv, w['dvalues'])
            assert np.array_equal(states.cpu().numpy(), w['states'])
            if not fuse:
                flat = oracle.concat_step_batches(ro.actions, want['returns'], ro.values, ro.log_probs)
                for got, full in zip(fields, flat):
                    assert np.array_equal(got.cpu().numpy(), full[w['idx']])


# ---------------------------------------------------------------------------------------------- rollout-time policy step
@pytest.mark.parametrize('kind,A', [('logits', 6), ('probs', 4), ('logits', 18)])
def test_policy_step_with_supplied_noise_vs_oracle(kind, A):
    rng = np.random.default_rng(A)
    n = 4097
    logits = rng.standard_normal((n, A)).astype(np.float32)
    actor = logits
    if kind == 'probs':
        e = np.exp(logits - logits.max(-1, keepdims=True))
        actor = (e / e.sum(-1, keepdims=True)).astype(np.float32)
    u = rng.random((n, A)).astype(np.float32).clip(1e-7, 1 - 1e-7)
    _, ent, lsm = oracle.categorical_logp_entropy(actor, np.zeros(n), kind == 'probs')
    want_a = np.argmax(lsm.astype(np.float64) - np.log(-np.log(u.astype(np.float64))), axis=-1)
    a, lp, en = ops.policy_step(cu(actor), actor_kind=kind, noise=cu(u))
    a = a.cpu().numpy()
    gap = np.sort(lsm.astype(np.float64) - np.log(-np.log(u.astype(np.float64))), axis=-1)
    clear = (gap[:, -1] - gap[:, -2]) > 1e-4                   # ignore numerically tied Gumbel scores
    assert np.array_equal(a[clear], want_a[clear].astype(np.float32)) and clear.mean() > 0.99
    close(lp, lsm[np.arange(n), a.astype(np.int64)])
    close(en, ent)


def test_policy_step_philox_sampler_statistics():
    logits = torch.tensor([[2.0, 0.0, -1.0, 0.5, 0.5, -3.0]], device=DEV).repeat(200_000, 1).contiguous()
    p = torch.softmax(logits[0], -1).cpu().numpy()
    a1, lp, en = ops.policy_step(logits, seed=7, offset=0)
    a2, _, _ = ops.policy_step(logits, seed=7, offset=0)
    a3, _, _ = ops.policy_step(logits, seed=7, offset=12)
    assert torch.equal(a1, a2) and not torch.equal(a1, a3)     # counter-based: reproducible, offset advances it
    freq = np.bincount(a1.cpu().numpy().astype(np.int64), minlength=6) / len(a1)
    sigma = np.sqrt(p * (1 - p) / len(a1))
    assert np.all(np.abs(freq - p) < 5 * sigma), (freq, p)
    assert abs(float(lp.mean()) - float((p * np.log(p)).sum())) < 0.01           # E[log p(a)] = -H
    loc = torch.zeros((100_000, 3), device=DEV)
    act, lp, en = ops.policy_step(loc, actor_kind='normal', seed=3)
    assert abs(float(act.mean())) < 0.02 and abs(float(act.var()) - 1.0) < 0.02
    close(lp, (-0.5 * (act.double() ** 2).sum(-1) - 1.5 * np.log(2 * np.pi)).cpu().numpy())
    assert abs(float(en[0]) - 1.5 * (1 + np.log(2 * np.pi))) < 1e-5


# ---------------------------------------------------------------------------------------------- next rows: ACER / TRPO reuse
def test_acer_retrace_bit_exact(golden):
    g = golden('acer_retrace')
    T, E = int(g['n_steps']), int(g['n_envs'])
    tm = lambda flat, steps=T: np.ascontiguousarray(flat.reshape(E, steps).T)
    values = tm(g['values'], T + 1)
    dones = np.concatenate([np.zeros((1, E), np.float32), tm(g['dones'])])
    got = ops.retrace_returns(cu(tm(g['rewards'])), cu(dones), cu(values[:-1]), cu(values[-1]), cu(tm(g['q_selected'])),
                              cu(tm(g['importance'])), float(g['gamma']))
    assert np.array_equal(got.cpu().numpy(), tm(g['returns']))
    rng = np.random.default_rng(9)
    for T, E in ((20, 4097), (1, 3), (333, 64)):
        f = lambda *s: rng.standard_normal(s).astype(np.float32)
        r, v, q, lv = f(T, E), f(T, E), f(T, E), f(E)
        d = (rng.random((T + 1, E)) < 0.1).astype(np.float32)
        imp = np.exp(f(T, E))
        want = oracle.retrace_returns(r, d, v, lv, q, imp, 0.99)
        got = ops.retrace_returns(cu(r), cu(d), cu(v), cu(lv), cu(q), cu(imp), 0.99)
        assert np.array_equal(got.cpu().numpy(), want)


def test_trpo_whole_batch_advantage_normalisation():
    """TRPO.train_step normalises over the WHOLE batch with no epsilon (xagents/trpo/agent.py:311-314):
    same kernels, one moments row, eps = 0."""
    ro = synthetic.make_rollout(64, 32, with_obs=False, epochs=0)
    ret = oracle.gae_returns(ro.rewards, ro.dones, ro.values, ro.last_values, 0.99, 0.95)
    adv = (ret - ro.values).reshape(-1)
    want = (adv - adv.mean()) / adv.std()
    got = ops.normalize_advantages(cu(ret.reshape(-1)), cu(ro.values.reshape(-1)), 0.0)
    close(got, want)


# ---------------------------------------------------------------------------------------------- small configs: graphs, A2C
def test_hotpath_cuda_graph_replay_equals_eager():
    """C1-shaped (CartPole: T=128, E=16, 4 fp32 features, 2 actions): one graph replay == the eager step."""
    from xagents_b200.hotpath import PPOHotPath
    T, E, A = 128, 16, 2
    ro = synthetic.make_rollout(T, E, obs_shape=(4,), obs_dtype='float32', n_actions=A, epochs=4, p_done=0.02)
    want = oracle.ppo_train_step(ro.obs, ro.rewards, ro.dones, ro.values, ro.last_values, ro.actions, ro.log_probs,
                                 ro.permutations, ro.new_logits, ro.new_values, mini_batches=4)
    hp = PPOHotPath(T, E, (4,), A, obs_dtype=torch.float32, device=DEV)
    hp.load(ro)
    hp.perms.copy_(torch.as_tensor(np.stack(ro.permutations)))
    for i, w in enumerate(want['minibatches']):
        hp.actor_out[i].copy_(torch.as_tensor(ro.new_logits[w['idx']]))
        hp.critic_out[i].copy_(torch.as_tensor(ro.new_values[w['idx']]))
    hp.prepare()
    hp.run()
    torch.cuda.synchronize()
    eager = hp.scalars.clone()
    replay = hp.capture()
    hp.scalars.zero_()
    hp.returns.zero_()
    replay()
    torch.cuda.synchronize()
    assert torch.equal(hp.scalars, eager)
    close(hp.returns, want['returns'])
    for i, w in enumerate(want['minibatches']):
        scale = max(abs(float(w[k])) for k in ('loss', 'pg', 'vl', 'entropy'))
        assert abs(float(hp.scalars[i, 0]) - w['loss']) <= REL * scale
    hp.rewards.mul_(2.0)                                   # new data in the same buffers, replay again
    replay()
    torch.cuda.synchronize()
    assert not torch.equal(hp.scalars, eager)
    hp.run()                                               # and the eager path still works after the capture
    torch.cuda.synchronize()


@pytest.mark.parametrize('case', ['a2c_image', 'a2c_vector'])
def test_a2c_hotpath_vs_reference_golden(golden, case):
    from xagents_b200.hotpath import A2CHotPath
    g = golden(case)
    T, E = g['returns'].shape
    A = g['actor'].shape[1]
    hp = A2CHotPath(T, E, A, gamma=float(g['gamma']), entropy_coef=float(g['entropy_coef']),
                    value_loss_coef=float(g['value_loss_coef']), device=DEV, scan_mode='sequential')
    tm = lambda flat: np.ascontiguousarray(np.asarray(flat).reshape(E, T).T)
    hp.rewards.copy_(cu(g['rewards']))
    hp.dones.copy_(cu(g['dones']))
    hp.last_values.copy_(cu(g['next_values']))
    hp.values.copy_(cu(tm(g['flat_values'])))
    hp.actions.copy_(cu(tm(g['flat_actions'])))
    # model outputs arrive env-major in the fixture; the hot path consumes them in time-major sample order
    hp.actor_out.copy_(cu(g['actor'].reshape(E, T, A).transpose(1, 0, 2).reshape(T * E, A)))
    hp.critic_out.copy_(cu(tm(g['critic'])).reshape(-1))
    hp.run()
    torch.cuda.synchronize()
    assert np.array_equal(hp.returns.cpu().numpy(), g['returns'])
    ref = float(g['loss'][0])
    assert abs(float(hp.scalars[0]) - ref) <= REL * max(abs(ref), abs(float(hp.scalars[2])), abs(float(hp.scalars[3])))
