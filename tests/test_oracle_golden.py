"""The oracle against fixtures produced by the reference's own code (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import oracle

PPO_CASES = ['ppo_image', 'ppo_cartpole', 'ppo_ragged', 'ppo_single_env', 'ppo_box', 'ppo_softmax']
A2C_CASES = ['a2c_image', 'a2c_vector', 'a2c_box', 'a2c_softmax']


def actor_kind(g):
    """Which branch of A2C.get_distribution produced the fixture (a2c/agent.py:50-63); older fixtures are all logits."""
    return str(g['actor_kind']) if 'actor_kind' in g.files else 'logits'


def logp_entropy(g, actor, actions):
    kind = actor_kind(g)
    if kind == 'normal':
        return oracle.diag_normal_logp_entropy(actor, actions.reshape(actor.shape))[:2]
    return oracle.categorical_logp_entropy(actor, actions.reshape(-1), kind == 'probs')[:2]


def test_kat_returns_bit_exact(golden):
    g = golden('kat_returns')
    got = oracle.gae_returns(g['rewards'], g['dones'], g['values'], g['next_values'], float(g['gamma']), float(g['lam']))
    assert got.dtype == np.float32 and np.array_equal(got, g['ppo_returns'])
    got = oracle.nstep_returns(g['rewards'], g['dones'], g['next_values'], float(g['gamma']))
    assert np.array_equal(got, g['a2c_returns'])
    flat_r, flat_v = oracle.concat_step_batches(g['ppo_returns'], g['values'])
    assert np.array_equal(flat_r, g['flat_returns']) and np.array_equal(flat_v, g['flat_values'])
    # SURVEY.md 8c KAT-1 / KAT-2 literal values
    np.testing.assert_allclose(g['ppo_returns'][0], [1.8897848, 0., -1.2676848], rtol=0, atol=1e-7)
    np.testing.assert_allclose(g['a2c_returns'][0], [1.9801, 0., -1.465498], rtol=0, atol=1e-7)


@pytest.mark.parametrize('case', PPO_CASES)
def test_ppo_returns_and_flatten_bit_exact(golden, case):
    g = golden(case)
    T, E = int(g['n_steps']), int(g['n_envs'])
    values = g['values'].reshape(T, E)
    got = oracle.gae_returns(g['rewards'], g['dones'], values, g['next_values'], float(g['gamma']), float(g['lam']))
    assert np.array_equal(got, g['returns'].reshape(T, E))
    flat = oracle.concat_step_batches(g['states_time_major'], got, values)
    assert np.array_equal(flat[0], g['flat_states'])
    assert np.array_equal(flat[1].reshape(-1), g['flat_returns'].reshape(-1))
    assert np.array_equal(flat[2].reshape(-1), g['flat_values'].reshape(-1))
    b = np.arange(T * E)
    rows = oracle.env_major_to_time_major(b, T, E)
    assert np.array_equal(g['states_time_major'].reshape((T * E,) + g['flat_states'].shape[1:])[rows], g['flat_states'])
    assert int(g['steps_after']) == T * E


@pytest.mark.parametrize('case', PPO_CASES)
def test_ppo_minibatches_and_loss(golden, case):
    g = golden(case)
    N = int(g['n_steps']) * int(g['n_envs'])
    B = int(g['mini_batch_size'])
    fields = [g['flat_states'], g['flat_actions'], g['flat_returns'], g['flat_values'], g['flat_log_probs']]
    mbs = oracle.gather_minibatches(fields, g['shuffles'], B)
    assert len(mbs) == len(g['losses']) == int(g['ppo_epochs']) * -(-N // B)
    for i, mb in enumerate(mbs):
        assert np.array_equal(mb[0], g[f'mb{i}_states'])
        r, v, lp, a = mb[2].reshape(-1), mb[3].reshape(-1), mb[4].reshape(-1), mb[1]
        assert np.array_equal(r, g[f'mb{i}_returns'].reshape(-1))
        adv = oracle.normalize_advantages(r, v, float(g['advantage_epsilon']))
        np.testing.assert_allclose(adv, g[f'mb{i}_advantages'].reshape(-1), rtol=1e-6, atol=1e-6)
        logp, ent = logp_entropy(g, g[f'mb{i}_actor'], a)
        sc = oracle.ppo_loss(logp, g[f'mb{i}_critic'].reshape(-1), ent, v, r, lp, adv, float(g['clip_norm']),
                             float(g['entropy_coef']), float(g['value_loss_coef']))
        ref_loss = g['losses'][i]
        assert abs(sc['loss'] - ref_loss) <= 1e-5 * abs(ref_loss), (i, sc['loss'], ref_loss)
        _, ref_ent, ref_vl, ref_pg = g['means'][i]
        scale = abs(ref_loss)
        assert abs(sc['entropy'] - ref_ent) <= 1e-5 * scale
        assert abs(sc['vl'] - 0.5 * ref_vl) <= 1e-5 * scale
        assert abs(sc['pg'] - ref_pg) <= 1e-5 * scale


@pytest.mark.parametrize('case', A2C_CASES)
def test_a2c_path(golden, case):
    g = golden(case)
    got = oracle.nstep_returns(g['rewards'], g['dones'], g['next_values'], float(g['gamma']))
    assert np.array_equal(got, g['returns'])
    assert np.array_equal(oracle.concat_step_batches(got)[0].reshape(-1), g['flat_returns'].reshape(-1))
    logp, ent = logp_entropy(g, g['actor'], g['flat_actions'])
    sc = oracle.a2c_loss(logp, g['critic'].reshape(-1), ent, g['flat_values'].reshape(-1),
                         g['flat_returns'].reshape(-1), float(g['entropy_coef']), float(g['value_loss_coef']))
    assert abs(sc['loss'] - g['loss'][0]) <= 1e-5 * abs(g['loss'][0])


def test_kat3_kat4_loss_values(golden):
    """SURVEY.md 8c KAT-3 / KAT-4 (hand-derivable), forward + closed-form backward."""
    g = golden('kat_returns')
    idx = np.array([7, 2, 9, 0, 5, 11])
    r, v = g['flat_returns'].reshape(-1)[idx], g['flat_values'].reshape(-1)[idx]
    logits = np.float32([[.1, -.2, .3], [1, 0, -1], [.5, .5, .5], [-.3, .8, .2], [2, -1, 0], [0, .1, -.1]])
    new_v = np.float32([.35, .6, -.45, .55, .25, .1])
    actions = np.array([0, 2, 1, 1, 0, 2])
    old_logp = np.float32([-1, -2.3, -1.2, -.7, -.3, -1.05])
    adv = oracle.normalize_advantages(r, v, 1e-8)
    np.testing.assert_allclose(adv, [-0.09828994, 0.45400777, -0.5042908, 0.6545685, 1.3239344, -1.8299301], atol=2e-6)
    logp, ent, _ = oracle.categorical_logp_entropy(logits, actions)
    sc = oracle.ppo_loss(logp, new_v, ent, v, r, old_logp, adv, .1, .01, .5)
    np.testing.assert_allclose([sc['pg'], sc['vl'], sc['entropy'], sc['loss']],
                               [-0.044882495, 1.152010560, 0.938469231, 0.521738112], atol=2e-6)
    dl, dv = oracle.ppo_loss_grads(logits, new_v, actions, v, r, old_logp, adv, .1, .01, .5)
    np.testing.assert_allclose(dv, [0.045666665, 0, 0, -0.11164874, -0.20891115, 0.175], atol=1e-6)
    np.testing.assert_allclose(dl[0], [9.9536581e-03, -3.8876350e-03, -6.0660210e-03], atol=1e-7)
    sc = oracle.a2c_loss(logp, new_v, ent, v, r, .01, .5)
    np.testing.assert_allclose([sc['pg'], sc['vl'], sc['loss']], [0.216375038, 2.162899017, 1.288439870], atol=2e-6)


@pytest.mark.parametrize('is_probs', [False, True])
def test_closed_form_grads_match_autograd(is_probs):
    from oracle import torch_ref
    rng = np.random.default_rng(5)
    n, A = 257, 6
    logits = rng.standard_normal((n, A)).astype(np.float32)
    actor = np.exp(logits - logits.max(-1, keepdims=True))
    actor = (actor / actor.sum(-1, keepdims=True)).astype(np.float32) if is_probs else logits
    actions = rng.integers(0, A, n)
    old_v = rng.standard_normal(n).astype(np.float32)
    new_v = (old_v + 0.15 * rng.standard_normal(n)).astype(np.float32)
    ret = rng.standard_normal(n).astype(np.float32)
    logp, _, _ = oracle.categorical_logp_entropy(actor, actions, is_probs)
    old_logp = (logp + 0.15 * rng.standard_normal(n)).astype(np.float32)
    adv = oracle.normalize_advantages(ret, old_v, 1e-8)
    sc, da, dv = torch_ref.ppo_loss_autograd(actor, new_v, actions, old_v, ret, old_logp, adv, .1, .01, .5, is_probs)
    dl, dvv = oracle.ppo_loss_grads(actor, new_v, actions, old_v, ret, old_logp, adv, .1, .01, .5, is_probs)
    np.testing.assert_allclose(dl, da, atol=1e-5 * np.abs(da).max())
    np.testing.assert_allclose(dvv, dv, atol=1e-5 * np.abs(dv).max())
    assert (dv == 0).any()                                           # saturated value-clip branch is exercised
    sc, da, dv = torch_ref.a2c_loss_autograd(actor, new_v, actions, old_v, ret, .01, .5, is_probs)
    dl, dvv = oracle.a2c_loss_grads(actor, new_v, actions, old_v, ret, .01, .5, is_probs)
    np.testing.assert_allclose(dl, da, atol=1e-5 * np.abs(da).max())
    np.testing.assert_allclose(dvv, dv, atol=1e-5 * np.abs(dv).max())


def test_acer_retrace_bit_exact_vs_reference_run(golden):
    """ACER.calculate_returns run by the reference itself (flat env-major in/out) vs the time-major oracle."""
    g = golden('acer_retrace')
    T, E = int(g['n_steps']), int(g['n_envs'])
    tm = lambda flat, steps=T: np.ascontiguousarray(flat.reshape(E, steps).T)
    values = tm(g['values'], T + 1)
    dones = np.concatenate([np.zeros((1, E), np.float32), tm(g['dones'])])     # row t+1 = flag after step t
    got = oracle.retrace_returns(tm(g['rewards']), dones, values[:-1], values[-1], tm(g['q_selected']), tm(g['importance']),
                                 float(g['gamma']))
    assert np.array_equal(got, tm(g['returns']))
