"""ACER on the hot-path kernels (xagents/acer/agent.py).  Retrace itself is pinned bit-exactly to the
reference's own `calculate_returns` in test_gpu_parity.py (golden acer_retrace.npz); here the rest of the update is checked
against an fp64 restatement of the reference's arithmetic in the reference's own env-major layout
(tests/acer_restatement.py), the device trajectory ring against plain indexing, and the train-step bookkeeping."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REL = 1e-5


def _golden_module():
    spec = importlib.util.spec_from_file_location('make_golden', os.path.join(HERE, 'golden', 'make_golden.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class TwoHeads(torch.nn.Module):
    def __init__(self, n_in, n_actions):
        super().__init__()
        self.trunk, self.policy, self.q = torch.nn.Linear(n_in, 16), torch.nn.Linear(16, n_actions), torch.nn.Linear(16, n_actions)

    def forward(self, x):
        h = torch.tanh(self.trunk(x.reshape(x.shape[0], -1)))
        return torch.softmax(self.policy(h), -1), self.q(h)


def _agent(T=6, E=4, A=3, obs=(5,), image=False, seed=0, slots=3, initial=2, **kw):
    from xagents_b200.agents import ACER, TorchModel
    from xagents_b200.buffers import create_buffers
    mg = _golden_module()
    rng = np.random.default_rng(seed)
    obs_s, rewards, dones, resets = mg._streams(rng, 12 * T, E, obs, image, 0.2)
    envs = [mg.ReplayEnv(obs_s[i], rewards[i], dones[i], resets[i], mg.Discrete(A)) for i in range(E)]
    torch.manual_seed(seed + 1)
    net = TorchModel(TwoHeads(int(np.prod(obs)), A).cuda(), output_is_softmax=True)
    buffers = create_buffers('acer', slots * E, 32, E, initial * E)
    return ACER(envs, net, buffers, n_steps=T, quiet=True, seed=seed + 2, **kw), net


def test_constructor_surface_buffers_and_errors():
    from xagents_b200.agents import ACER
    from xagents_b200.buffers import ReplayBuffer1, create_buffers
    agent, net = _agent()
    for name, want in (('ema_alpha', 0.99), ('replay_ratio', 4), ('epsilon', 1e-6), ('importance_c', 10.0), ('delta', 1),
                       ('trust_region', True), ('entropy_coef', 0.01), ('value_loss_coef', 0.5), ('grad_norm', 0.5)):
        assert getattr(agent, name) == want
    assert agent.batch_dtypes == ['uint8', 'float32', 'int32', 'float32', 'float32'] and agent.actor_kind == 'probs'
    assert agent.batch_shapes == [(4 * 7, 5), (24,), (24,), (24,), (24, 3)]
    assert (agent.buffers[0].size, agent.buffers[0].initial_size, agent.buffers[0].batch_size) == (3, 2, 1)
    with pytest.raises(AssertionError, match='Buffer batch size should be 1 for ACER, got 2'):
        ACER(agent.envs, net, [ReplayBuffer1(4, batch_size=2) for _ in range(4)], n_steps=6, quiet=True)
    with pytest.raises(AssertionError, match='off-policy'):
        create_buffers('dqn', 100, 32, 4)


def test_device_trajectory_ring_replays_exactly_what_the_rollouts_wrote():
    agent, _ = _agent(T=5, E=3, obs=(6, 6, 2), image=True, slots=3)
    kept = []
    for _ in range(5):                                             # 5 rollouts into 3 slots: the two oldest are overwritten
        kept.append([f.clone() for f in agent.get_batch()])
    assert agent.ring.current_size == 3 and agent.buffers[1].current_size == 3 and agent.ring.states.dtype == torch.uint8
    by_slot = {i % 3: kept[i] for i in range(5)}                   # slot -> latest trajectory written there
    slots = np.array([2, 0, 1])
    got = agent.ring.gather(slots)
    for f, field in enumerate(got):
        for e, s in enumerate(slots):
            assert torch.equal(field[:, e], by_slot[int(s)][f][:, e])
    sampled = agent.concat_buffer_samples()
    assert [tuple(x.shape) for x in sampled] == [(6, 3, 6, 6, 2), (5, 3), (5, 3), (5, 3), (5, 3, agent.n_actions)]


@pytest.mark.parametrize('trust_region', [True, False])
def test_update_gradients_match_the_fp64_restatement(trust_region):
    from acer_restatement import acer_output_gradients
    agent, net = _agent(T=8, E=5, A=4, trust_region=trust_region, delta=0.2, grad_norm=None)
    states, rewards, actions, dones, prev = [f.clone() for f in agent.get_batch()]
    T, E, A = 8, 5, 4
    with torch.no_grad():
        probs, q = net.module(states.reshape(-1, 5))
    avg = torch.softmax(torch.randn((T + 1) * E, A, device='cuda', generator=torch.Generator('cuda').manual_seed(4)), -1)
    agent._avg_outputs = lambda s: avg                             # an averaged policy that differs from the current one
    want_p, want_q, want_r = acer_output_gradients(
        probs.view(T + 1, E, A), q.view(T + 1, E, A), avg.view(T + 1, E, A), rewards, dones, actions, prev, gamma=agent.gamma,
        epsilon=agent.epsilon, importance_c=agent.importance_c, delta=agent.delta, trust_region=trust_region,
        entropy_coef=agent.entropy_coef, value_loss_coef=agent.value_loss_coef)
    seen = {}
    step = net.backward_and_step
    net.backward_and_step = lambda da, dq, gn: (seen.update(p=da.clone(), q=dq.clone(), clip=gn), step(da, dq, gn))
    agent.update_gradients(states, rewards, actions, dones, prev)
    for got, want in ((seen['p'], want_p), (seen['q'], want_q)):
        got = got.view(T + 1, E, A).double()
        assert (got - want).abs().max() <= REL * want.abs().max()
        assert torch.count_nonzero(got[T]) == 0                    # clip_last_step: the bootstrap row gets no gradient
    if trust_region:                                               # the projection must actually have acted in this case
        free_p, _, _ = acer_output_gradients(
            probs.view(T + 1, E, A), q.view(T + 1, E, A), avg.view(T + 1, E, A), rewards, dones, actions, prev, gamma=agent.gamma,
            epsilon=agent.epsilon, importance_c=agent.importance_c, delta=agent.delta, trust_region=False,
            entropy_coef=agent.entropy_coef, value_loss_coef=agent.value_loss_coef)
        assert (free_p - want_p).abs().max() > 1e-3 * want_p.abs().max()
    assert seen['clip'] is None and net.step == 1


def test_train_step_bookkeeping_replay_and_averaged_network():
    agent, net = _agent(T=6, E=4, slots=3, initial=2, replay_ratio=3, ema_alpha=0.9)
    start = net.flat_param.clone()
    assert torch.equal(agent.avg_flat, start)
    agent.train_step()                                             # buffer below its initial size: on-policy update only
    assert (agent.buffer_current_size, net.step, agent.steps) == (1, 1, 24)
    assert torch.equal(agent.avg_flat, net.flat_param)             # EMA shadow starts as a copy at the first apply
    before_avg, updates = agent.avg_flat.clone(), net.step
    agent._poisson = type('Fixed', (), {'poisson': staticmethod(lambda lam: 2)})()
    agent.train_step()                                             # initial size reached: 1 on-policy + 2 replay updates
    assert (agent.buffer_current_size, net.step - updates, agent.steps) == (2, 3, 48)
    assert not torch.equal(agent.avg_flat, before_avg) and not torch.equal(agent.avg_flat, net.flat_param)
    # one more explicit EMA step follows the definition shadow -= (1 - decay) * (shadow - value)
    shadow = agent.avg_flat.clone()
    agent.update_avg_weights()
    assert torch.allclose(agent.avg_flat, shadow - (1 - 0.9) * (shadow - net.flat_param), rtol=1e-6, atol=1e-8)
    assert torch.isfinite(net.flat_param).all()


@pytest.mark.timeout(300)
def test_cli_train_acer(tmp_path):
    """`xagents train acer`: default cnn `.cfg` (softmax policy head + n_actions-wide critic) on synthetic frames, and a
    user `.cfg` on CartPole (the reference ships no ann default for ACER: asserted like common.py:455-458)."""
    from xagents_b200 import cli
    ex = cli.Executor()
    ex.execute(['train', 'acer', '--env', 'SyntheticAtari-v0', '--n-envs', '4', '--n-steps', '8', '--max-steps', '160', '--quiet', '--seed', '3',
                '--buffer-max-size', '16', '--buffer-initial-size', '8', '--conv-dims', '2', '--trust-region'])
    agent = ex.agent
    assert type(agent).__name__ == 'ACER' and agent.trust_region and agent.grad_norm == 10 and agent.ring.slots == 4
    assert agent.net.output_is_softmax and agent.net.n_params == 1_687_719 - (512 + 1) + (512 * 6 + 6)
    assert agent.steps == 160 and agent.buffer_current_size == 5 and agent.net.step >= 5 and torch.isfinite(agent.net.flat_param).all()
    with pytest.raises(AssertionError, match='You should specify `model_cfg`. No default ANN model found'):
        cli.execute(['train', 'acer', '--env', 'CartPole-v1', '--max-steps', '10', '--quiet'])
    cfg = tmp_path / 'ann-actor-critic.cfg'
    cfg.write_text('[dense-0]\nunits=64\nactivation=tanh\ninitializer=orthogonal\ngain=1.4142135\ncommon=1\n'
                   '[dense-1]\nactivation=softmax\ninitializer=orthogonal\ngain=0.01\noutput=1\n'
                   '[dense-2]\ninitializer=orthogonal\ngain=1.0\noutput=1\n')
    ex = cli.Executor()
    ex.execute(['train', 'acer', '--env', 'CartPole-v1', '--n-envs', '8', '--max-steps', '16000', '--quiet', '--seed', '1', '--model', str(cfg),
                '--buffer-max-size', '800', '--buffer-initial-size', '80'])
    agent = ex.agent
    assert agent.n_steps == 20 and not agent.trust_region and agent.steps >= 16000
    assert agent.net.step > agent.buffer_current_size               # replay updates happened on top of the on-policy ones
    agent.update_metrics()
    assert torch.isfinite(agent.net.flat_param).all() and agent.games > 100 and np.isfinite(agent.mean_reward)


@pytest.mark.parametrize('case', ['acer_update_trust_region', 'acer_update_plain'])
def test_update_gradients_vs_the_reference_run(golden, case):
    """The whole ACER.update_gradients on the device -- values, clip_last_step, importance weights, the Retrace kernel, losses,
    trust-region projection -- against the fixture made by the REFERENCE'S OWN update_gradients (acer/agent.py:262-339; float64
    shim, tape answered by central differences of the reference's calculate_losses: tests/golden/make_golden.py).  The network is
    replaced by the fixture's model outputs; layouts are converted env-major (reference) -> time-major (this repo)."""
    g = golden(case)
    T, E, A = int(g['n_steps']), int(g['n_envs']), int(g['n_actions'])
    n = T * E
    agent, net = _agent(T=T, E=E, A=A, trust_region=bool(g['trust_region']), delta=float(g['delta']), grad_norm=None)
    assert (agent.epsilon, agent.importance_c, agent.entropy_coef, agent.value_loss_coef, agent.gamma) == (
        float(g['epsilon']), float(g['importance_c']), float(g['entropy_coef']), float(g['value_loss_coef']), float(g['gamma']))
    cu = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32).cuda()
    full_tm = lambda x: cu(x.reshape(E, T + 1, A).swapaxes(0, 1).reshape((T + 1) * E, A))        # [E(T+1), A] env-major -> [(T+1)E, A]
    steps_tm = lambda x: cu(x.reshape((E, T) + x.shape[1:]).swapaxes(0, 1))                    # [E T, ...] -> [T, E, ...]
    probs_full, q_full, avg_full = full_tm(g['full_action_probs']), full_tm(g['full_critic_logits']), full_tm(g['full_avg_action_probs'])
    net.forward = lambda states, training=True: (probs_full, q_full)
    agent._avg_outputs = lambda states: avg_full
    seen = {}
    net.backward_and_step = lambda da, dq, gn: seen.update(p=da.clone(), q=dq.clone())
    states = torch.zeros((T + 1, E) + agent.input_shape, dtype=agent.ring.states.dtype if hasattr(agent, 'ring') else torch.float32, device='cuda')
    agent.update_gradients(states, steps_tm(g['rewards']), steps_tm(g['actions'].astype(np.float32)), steps_tm(g['dones']),
                           steps_tm(g['previous_action_probs']))
    torch.cuda.synchronize()
    to_env_major = lambda x: x[:n].view(T, E, A).transpose(0, 1).reshape(n, A).double().cpu().numpy()
    want_p = g['output_grads'] if bool(g['trust_region']) else g['d_loss_d_action_probs']
    want_q = g['d_value_loss_d_critic_logits'] if bool(g['trust_region']) else g['d_loss_d_critic_logits']
    for got, want in ((to_env_major(seen['p']), want_p), (to_env_major(seen['q']), want_q)):
        assert np.abs(got - want).max() <= REL * np.abs(want).max()
    assert torch.count_nonzero(seen['p'][n:]) == 0 and torch.count_nonzero(seen['q'][n:]) == 0     # clip_last_step
    losses = agent.last_losses
    if bool(g['trust_region']):
        assert abs(float(losses[0]) - float(g['loss'])) <= REL * abs(float(g['loss']))
        assert abs(float(losses[1]) - float(g['value_loss'])) <= REL * abs(float(g['value_loss']))
    else:
        assert abs(float(losses) - float(g['loss'])) <= REL * max(abs(float(g['loss'])), 1.0)
