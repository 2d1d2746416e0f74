"""Size-independent properties of the oracle (the checker itself gets checked): hypothesis draws shapes, done patterns and
hyper-parameters; the properties are the ones the GPU parity tests rely on at sizes the fixtures do not reach."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle

SHAPES = st.tuples(st.integers(1, 40), st.integers(1, 9))
SEEDS = st.integers(0, 2 ** 31 - 1)


def _rollout(seed, T, E, p_done):
    rng = np.random.default_rng(seed)
    f = lambda *s: rng.standard_normal(s).astype(np.float32)
    return f(T, E), f(T, E), f(E), (rng.random((T + 1, E)) < p_done).astype(np.float32), rng


@settings(max_examples=60, deadline=None)
@given(SHAPES, SEEDS, st.sampled_from([0.0, 0.05, 0.5, 1.0]), st.floats(0.5, 1.0), st.floats(0.0, 1.0))
def test_gae_fp32_tracks_its_fp64_twin_and_cuts_at_dones(shape, seed, p_done, gamma, lam):
    T, E = shape
    rewards, values, last, dones, rng = _rollout(seed, T, E, p_done)
    got = oracle.gae_returns(rewards, dones, values, last, gamma, lam)
    twin = oracle.gae_returns(rewards, dones, values, last, gamma, lam, dtype=np.float64)
    assert got.dtype == np.float32 and got.shape == (T, E)
    assert np.abs(got - twin).max() <= 1e-5 * max(np.abs(twin).max(), 1.0)
    if p_done == 1.0:                                              # every step terminal: returns are the rewards
        np.testing.assert_allclose(got, rewards, atol=1e-6)
    # a done after step t cuts the recurrence: nothing later than t can influence returns[:t+1] of that environment
    t, e = int(rng.integers(T)), int(rng.integers(E))
    cut = dones.copy()
    cut[t + 1, e] = 1.0
    base = oracle.gae_returns(rewards, cut, values, last, gamma, lam)
    r2, v2, l2 = rewards.copy(), values.copy(), last.copy()
    r2[t + 1:, e] += 3.0
    v2[t + 1:, e] -= 2.0
    l2[e] += 5.0
    moved = oracle.gae_returns(r2, cut, v2, l2, gamma, lam)
    assert np.array_equal(base[:t + 1, e], moved[:t + 1, e])
    others = [k for k in range(E) if k != e]
    assert np.array_equal(base[:, others], moved[:, others])      # environments never mix


@settings(max_examples=40, deadline=None)
@given(SHAPES, SEEDS, st.floats(0.5, 1.0))
def test_returns_family_relations(shape, seed, gamma):
    T, E = shape
    rewards, values, last, dones, rng = _rollout(seed, T, E, 0.1)
    nstep = oracle.nstep_returns(rewards, dones, last, gamma)
    # GAE with lam = 1 telescopes to the n-step return (up to fp32 re-association)
    gae1 = oracle.gae_returns(rewards, dones, values, last, gamma, 1.0, dtype=np.float64)
    nstep64 = oracle.nstep_returns(rewards, dones, last, gamma, dtype=np.float64)
    np.testing.assert_allclose(gae1, nstep64, atol=1e-9 * max(1.0, np.abs(nstep64).max()))
    # Retrace (acer/agent.py:198-208): with truncated weights min(1, rho) == 1 and q(s, a) == V(s) the correction term
    # vanishes and the target is the n-step return; with rho == 0 it collapses to the one-step TD target
    heavy = (1.0 + np.abs(rng.standard_normal((T, E)))).astype(np.float32)
    retrace = oracle.retrace_returns(rewards, dones, values, last, values, heavy, gamma)
    np.testing.assert_allclose(retrace, nstep, atol=2e-5 * max(1.0, np.abs(nstep).max()))
    one_step = oracle.retrace_returns(rewards, dones, values, last, values, np.zeros((T, E), np.float32), gamma)
    nxt = np.concatenate([values[1:], last[None]])
    np.testing.assert_allclose(one_step, rewards + np.float32(gamma) * nxt * (1 - dones[1:]), atol=1e-5 * max(1.0, np.abs(nxt).max()))
    # gamma = 1, no dones: suffix sums + bootstrap
    free = oracle.nstep_returns(rewards, np.zeros_like(dones), last, 1.0, dtype=np.float64)
    want = np.cumsum(rewards[::-1].astype(np.float64), axis=0)[::-1] + last
    np.testing.assert_allclose(free, want, atol=1e-9)


@settings(max_examples=40, deadline=None)
@given(SHAPES, SEEDS, st.integers(1, 7), st.integers(1, 3))
def test_flatten_and_minibatches_are_pure_index_work(shape, seed, mini_batches, epochs):
    T, E = shape
    N = T * E
    rng = np.random.default_rng(seed)
    obs = rng.integers(0, 256, (T, E, 3, 2), dtype=np.uint8)
    tag = np.arange(N, dtype=np.float32).reshape(T, E)            # time-major sample number
    flat_obs, flat_tag = oracle.concat_step_batches(obs, tag)
    b = np.arange(N)
    rows = oracle.env_major_to_time_major(b, T, E)
    assert sorted(rows.tolist()) == list(range(N))                 # the remap is a bijection ...
    assert np.array_equal(flat_tag.reshape(-1), rows.astype(np.float32))          # ... and it is the flatten's order
    assert np.array_equal(flat_obs, obs.reshape((N, 3, 2))[rows])
    B = max(N // mini_batches, 1)
    perms = [rng.permutation(N).astype(np.int32) for _ in range(epochs)]
    mbs = oracle.gather_minibatches([flat_obs, flat_tag], perms, B)
    per_epoch = -(-N // B)
    assert len(mbs) == epochs * per_epoch
    assert [len(mb[0]) for mb in mbs[:per_epoch]] == [hi - lo for lo, hi in oracle.minibatch_slices(N, B)]
    for k in range(epochs):                                        # each epoch visits every sample exactly once
        seen = np.concatenate([mb[1].reshape(-1) for mb in mbs[k * per_epoch:(k + 1) * per_epoch]])
        assert sorted(seen.tolist()) == sorted(flat_tag.reshape(-1).tolist())
        first = mbs[k * per_epoch]
        assert np.array_equal(first[0], flat_obs[perms[k][:len(first[0])]])        # bytes move verbatim


@settings(max_examples=40, deadline=None)
@given(st.integers(2, 300), st.integers(2, 12), SEEDS, st.booleans())
def test_normalisation_and_loss_gradient_invariants(n, A, seed, is_probs):
    rng = np.random.default_rng(seed)
    ret, old_v = rng.standard_normal(n).astype(np.float32) * 3 + 1, rng.standard_normal(n).astype(np.float32)
    adv = oracle.normalize_advantages(ret, old_v, 1e-8)
    if np.std(ret - old_v) > 1e-3:
        assert abs(float(adv.mean(dtype=np.float64))) < 1e-4 and abs(float(adv.std(dtype=np.float64)) - 1.0) < 1e-3
    logits = rng.standard_normal((n, A)).astype(np.float32)
    probs = np.exp(logits - logits.max(-1, keepdims=True))
    probs = (probs / probs.sum(-1, keepdims=True)).astype(np.float32)
    actor = probs if is_probs else logits
    actions = rng.integers(0, A, n)
    new_v = (old_v + 0.2 * rng.standard_normal(n)).astype(np.float32)
    logp, ent, lsm = oracle.categorical_logp_entropy(actor, actions, is_probs)
    assert np.all(ent >= -1e-6) and np.all(ent <= np.log(A) + 1e-5) and np.all(logp <= 1e-6)
    np.testing.assert_allclose(np.exp(lsm).sum(-1), 1.0, atol=1e-5)
    old_lp = (logp + 0.2 * rng.standard_normal(n)).astype(np.float32)
    d_actor, d_v = oracle.ppo_loss_grads(actor, new_v, actions, old_v, ret, old_lp, adv, 0.1, 0.01, 0.5, is_probs)
    if not is_probs:                                               # softmax is shift-invariant: logit gradients sum to zero per row
        assert np.abs(d_actor.sum(-1)).max() <= 1e-6
    # the value gradient vanishes exactly where the clipped branch is selected and saturated
    c = np.float32(0.1)
    e1 = np.square(new_v - ret)
    e2 = np.square(old_v + np.clip(new_v - old_v, -c, c) - ret)
    saturated = (e2 > e1) & (np.abs(new_v - old_v) > c)
    assert np.all(d_v[saturated] == 0)
    sc = oracle.ppo_loss(logp, new_v, ent, old_v, ret, old_lp, adv, 0.1, 0.01, 0.5)
    assert np.isclose(sc['loss'], sc['pg'] - 0.01 * sc['entropy'] + 0.5 * sc['vl'], atol=1e-6)
