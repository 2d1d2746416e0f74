"""TRPO on the PPO hot path (xagents/trpo/agent.py).  The reference's TRPO cannot be replayed through
the NumPy shim that produced the PPO / A2C fixtures (nested GradientTapes), so its pieces are checked against their
definitions in fp64: whole-batch advantage normalisation, Fisher-vector products against the explicit Hessian of the KL
divergence, conjugate gradients against a direct solve, the trust-region conditions of the line search, and the critic
update count; plus an end-to-end run through the command line."""
import importlib.util
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _golden_module():
    spec = importlib.util.spec_from_file_location('make_golden', os.path.join(HERE, 'golden', 'make_golden.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _mlp(n_in, n_hidden, n_out, seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(n_in, n_hidden), torch.nn.Tanh(), torch.nn.Linear(n_hidden, n_out)).cuda()


def _agent(T=16, E=4, A=3, obs=(5,), seed=0, **kw):
    from xagents_b200.agents import TRPO
    mg = _golden_module()
    rng = np.random.default_rng(seed)
    obs_s, rewards, dones, resets = mg._streams(rng, 4 * T, E, obs, False, 0.1)
    envs = [mg.ReplayEnv(obs_s[i], rewards[i], dones[i], resets[i], mg.Discrete(A)) for i in range(E)]
    actor, critic = _mlp(obs[0], 8, A, seed + 1), _mlp(obs[0], 8, 1, seed + 2)
    return TRPO(envs, actor, critic, n_steps=T, quiet=True, seed=seed + 3, **kw)


def _flat_batch(agent):
    states, actions, returns, values, _ = agent.get_batch()
    return states, actions, returns, values


def test_constructor_surface_and_defaults():
    agent = _agent()
    for name, want in (('max_kl', 1e-3), ('cg_iterations', 10), ('cg_residual_tolerance', 1e-10), ('cg_damping', 1e-3),
                       ('actor_iterations', 10), ('critic_iterations', 3), ('fvp_n_steps', 5), ('lam', 0.95), ('ppo_epochs', 4),
                       ('mini_batches', 4), ('entropy_coef', 0.01)):
        assert getattr(agent, name) == want
    assert agent.output_models == [agent.actor, agent.critic] and agent.model is agent.actor
    assert agent.actor.role == 'actor' and agent.critic.role == 'critic'
    from xagents_b200.agents import TRPO, TorchModel
    with pytest.raises(AssertionError, match='role'):
        TRPO(agent.envs, TorchModel(_mlp(5, 8, 3, 0)), _mlp(5, 8, 1, 0), n_steps=4, quiet=True)


def test_whole_batch_advantages_and_fvp_subsample_follow_the_env_major_order():
    """trpo/agent.py:316-319: (adv - mean) / std over the whole env-major batch, no epsilon; states[::fvp_n_steps]."""
    from xagents_b200 import ops
    agent = _agent(T=12, E=5)
    states, actions, returns, values = _flat_batch(agent)
    T, E, n = 12, 5, 60
    ids = torch.arange(n, dtype=torch.int32, device='cuda')
    adv = ops.normalize_advantages(returns.tensor, values.tensor, 0.0, idx=ids, time_major=(T, E)).cpu().numpy()
    r = agent.ro_returns.cpu().numpy().swapaxes(0, 1).reshape(-1).astype(np.float64)
    v = agent.ro_values.cpu().numpy().swapaxes(0, 1).reshape(-1).astype(np.float64)
    want = ((r - v) - (r - v).mean()) / (r - v).std()
    assert np.abs(adv - want).max() <= 1e-5 * np.abs(want).max()
    sub = ops.gather_rows(states.tensor, ids[::5].contiguous(), time_major=(T, E)).cpu().numpy()
    flat = agent.ro_states.cpu().numpy().swapaxes(0, 1).reshape(n, -1)
    assert np.array_equal(sub, flat[::5])


def test_fisher_vector_product_and_conjugate_gradients_against_the_explicit_hessian():
    agent = _agent(cg_iterations=200, cg_residual_tolerance=1e-16, cg_damping=1e-2)
    states, *_ = _flat_batch(agent)
    agent.at_step_start()
    x = states.materialize()
    n = agent.actor.n_params
    theta0 = agent.actor.flat_param[:n].double().clone()
    x64 = x.double()

    def logits(theta, inp):                                        # the actor, functionally, in fp64
        w1, b1, w2, b2 = theta[:40].view(8, 5), theta[40:48], theta[48:72].view(3, 8), theta[72:75]
        return torch.tanh(inp @ w1.T + b1) @ w2.T + b2

    def kl(theta):
        lo = torch.log_softmax(logits(theta0, x64), -1)
        ln = torch.log_softmax(logits(theta, x64), -1)
        return (lo.exp() * (lo - ln)).sum(-1).mean()

    assert n == 75
    hessian = torch.autograd.functional.hessian(kl, theta0)        # = the Fisher matrix at old == new
    torch.manual_seed(5)
    tangent = torch.randn(n, device='cuda')
    got = agent.calculate_fvp(tangent, x).double()
    want = hessian @ tangent.double() + agent.cg_damping * tangent.double()
    assert (got - want).norm() <= 1e-4 * want.norm()
    kl_now, old, new = agent.calculate_kl_divergence(x)
    assert abs(float(kl_now)) < 1e-7 and torch.equal(old, new.detach())
    g = torch.randn(n, device='cuda')
    solved = agent.conjugate_gradients(g, x).double()
    direct = torch.linalg.solve(hessian + agent.cg_damping * torch.eye(n, dtype=torch.float64, device='cuda'), g.double())
    assert (solved - direct).norm() <= 2e-2 * direct.norm()


def test_train_step_respects_the_trust_region_and_updates_the_critic():
    agent = _agent(T=32, E=8, max_kl=1e-2, ppo_epochs=2, mini_batches=4, critic_iterations=3)
    actor_before, critic_before = agent.actor.flat_param.clone(), agent.critic.flat_param.clone()
    agent.at_step_start()
    agent.train_step()
    torch.cuda.synchronize()
    info = agent.last_update
    assert torch.equal(agent.old_actor_flat, actor_before)
    if info['accepted']:
        assert info['kl_after'] <= 1.5 * agent.max_kl and info['surrogate_after'] > info['surrogate_before']
        assert not torch.equal(actor_before, agent.actor.flat_param)
        assert info['step_scale'] in [0.5 ** k for k in range(agent.actor_iterations)]
    else:
        assert torch.equal(actor_before, agent.actor.flat_param)
    assert info['accepted'], 'with max_kl=1e-2 on a fresh policy the line search is expected to accept a step'
    assert agent.actor.step == 0                                   # the actor never goes through Adam
    assert agent.critic.step == 3 * 2 * 4 and not torch.equal(critic_before, agent.critic.flat_param)
    assert agent.steps == 32 * 8 and torch.isfinite(agent.actor.flat_param).all() and torch.isfinite(agent.critic.flat_param).all()
    # the accepted step lies along full_step: moving on changes old_actor_flat at the next step start
    agent.at_step_start()
    assert torch.equal(agent.old_actor_flat, agent.actor.flat_param)


@pytest.mark.timeout(300)
def test_cli_train_trpo_on_cartpole_and_on_synthetic_frames():
    """`xagents train trpo` with the default actor / critic `.cfg` files (ann for CartPole, cnn for frames)."""
    from xagents_b200 import cli
    ex = cli.Executor()
    ex.execute(['train', 'trpo', '--env', 'CartPole-v1', '--n-envs', '16', '--n-steps', '128', '--max-steps', '40960', '--seed', '1',
                '--max-kl', '0.01', '--quiet'])
    agent = ex.agent
    assert type(agent).__name__ == 'TRPO' and agent.lam == 1.0 and agent.entropy_coef == 0 and agent.steps >= 40960
    assert agent.actor.n_params == 4 * 64 + 64 + 64 * 64 + 64 + 64 * 2 + 2 and agent.critic.step == 20 * 3 * 16
    agent.update_metrics()
    assert agent.best_reward > 30.0, f'TRPO did not improve on CartPole: best mean reward {agent.best_reward}'
    ex = cli.Executor()
    ex.execute(['train', 'trpo', '--env', 'SyntheticAtari-v0', '--n-envs', '4', '--n-steps', '16', '--max-steps', '128', '--quiet'])
    agent = ex.agent
    assert agent.ro_states.dtype == torch.uint8 and agent.actor.n_params == 28224 * 128 + 128 + 128 * 6 + 6
    assert agent.steps == 128 and torch.isfinite(agent.actor.flat_param).all() and torch.isfinite(agent.critic.flat_param).all()


def test_losses_kl_and_critic_minibatches_vs_the_reference_run(golden):
    """Surrogate objective and mean KL(old || new) against the REFERENCE'S OWN TRPO.calculate_losses / calculate_kl_divergence
    (trpo/agent.py:179-224) on linear networks, and the critic's value loss per shuffled minibatch against the reference's own
    update_critic_weights (:279-297) -- fixture tests/golden/trpo_losses.npz (make_golden.py --updates-only)."""
    from xagents_b200.agents import TRPO
    g = golden('trpo_losses')
    T, E, A = int(g['n_steps']), int(g['n_envs']), int(g['n_actions'])
    n, f_dim = T * E, g['states'].shape[1]
    mg = _golden_module()
    rng = np.random.default_rng(0)
    obs_s, rewards, dones, resets = mg._streams(rng, 2 * T, E, (f_dim,), False, 0.1)
    envs = [mg.ReplayEnv(obs_s[i], rewards[i], dones[i], resets[i], mg.Discrete(A)) for i in range(E)]

    def linear(w):
        m = torch.nn.Linear(w.shape[0], w.shape[1], bias=False)
        with torch.no_grad():
            m.weight.copy_(torch.as_tensor(w.T))
        return m.cuda()

    agent = TRPO(envs, linear(g['w_actor']), linear(g['w_critic']), n_steps=T, quiet=True, entropy_coef=float(g['entropy_coef']),
                 critic_iterations=int(g['critic_iterations']), ppo_epochs=int(g['ppo_epochs']), mini_batches=int(g['mini_batches']))
    agent.old_actor_flat[:A * f_dim].copy_(torch.as_tensor(np.ascontiguousarray(g['w_old_actor'].T).reshape(-1)).cuda())   # the "old" actor
    cu = lambda x: torch.as_tensor(np.ascontiguousarray(x)).cuda()
    states, actions, advantages = cu(g['states']), cu(g['actions']), cu(g['advantages'])
    with torch.no_grad():
        surrogate, kl = agent.calculate_losses(states, actions, advantages)
    assert abs(float(surrogate) - float(g['surrogate_loss'])) <= 1e-5 * max(abs(float(g['surrogate_loss'])), float(np.abs(g['advantages']).mean()))
    assert abs(float(kl) - float(g['kl_divergence'])) <= 1e-4 * float(g['kl_divergence'])      # a difference of log-probabilities ~1e-3: fp32
    # whole-batch advantage normalisation (trpo/agent.py:316-319) as the reference's shim evaluated it
    raw = g['raw_advantages']
    from xagents_b200 import ops
    zeros = torch.zeros(n, device='cuda')
    got = ops.normalize_advantages(cu(raw), zeros, 0.0).cpu().numpy()
    assert np.abs(got - g['advantages']).max() <= 1e-5 * np.abs(g['advantages']).max()
    # the critic's minibatches: same shuffles -> same value losses, in the same order
    shuffles = [s for s in g['critic_shuffles']]
    agent.permutation_source = lambda epoch, it=iter(shuffles): next(it)
    returns = cu(g['returns'])
    losses = []
    for _ in range(agent.critic_iterations):
        for states_mb, returns_mb in agent.get_mini_batches(states, returns):
            _, values = agent.critic.forward(states_mb, training=False)
            losses.append(float(((values - returns_mb) ** 2).mean()))
    want = g['critic_value_losses']
    assert len(losses) == len(want) and np.abs(np.asarray(losses) - want).max() <= 1e-5 * np.abs(want).max()
