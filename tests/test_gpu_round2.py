"""Round-2 parity cases: the C3-sized prepared pipeline against the oracle, the agents on the prepared pipelines, rollouts
fed from host memory, devices other than cuda:0, the zero-probability branch of Categorical(probs=), the debug index
check, and the fused peer-memory gradient all-reduce + Adam (2 GPUs)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

import oracle
from xagents_b200 import ops, synthetic

pytestmark = pytest.mark.gpu
REL = 1e-5
DEV = 'cuda:0'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.timeout(600)
def test_hotpath_at_c3_size_vs_oracle():
    """BASELINE config C3 itself (n_envs=256, n_steps=128, 4 epochs x 4 minibatches of 8192 frames, default launch
    schedule: one gather launch with per-minibatch completion counters) -- the exact object bench.py times -- against oracle.ppo_train_step: returns bit-exact, every
    minibatch's loss / pg / value / entropy <= 1e-5, three sampled minibatches' gathered frames bit-exact."""
    from xagents_b200.hotpath import PPOHotPath
    T, E, K, M = 128, 256, 4, 4
    ro = synthetic.make_rollout(T, E, obs_shape=(84, 84, 4), epochs=K, p_done=0.01)
    want = oracle.ppo_train_step(ro.obs, ro.rewards, ro.dones, ro.values, ro.last_values, ro.actions, ro.log_probs,
                                 ro.permutations, ro.new_logits, ro.new_values, mini_batches=M, keep_states=False)
    hp = PPOHotPath(T, E, (84, 84, 4), ro.n_actions, ppo_epochs=K, mini_batches=M, device=DEV, scan_mode='sequential')
    assert hp.sync == 'progress' and hp.group_sizes == [16] and hp.B == 8192 and hp.n_mb == 16      # one launch, per-minibatch counters
    hp.load(ro)
    hp.perms.copy_(torch.as_tensor(np.stack(ro.permutations)))
    for i, w in enumerate(want['minibatches']):
        hp.actor_out[i].copy_(torch.as_tensor(ro.new_logits[w['idx']]))
        hp.critic_out[i].copy_(torch.as_tensor(ro.new_values[w['idx']]))
    hp.prepare()
    sampled, kept = (0, 9, 15), {}

    def after_loss(i):
        if i in sampled:
            kept[i] = hp.minibatch_views(i)[0].clone()

    hp.run(after_loss=after_loss)
    torch.cuda.synchronize()
    assert np.array_equal(hp.returns.cpu().numpy(), want['returns'])          # sequential scan: the reference's operation order
    sc = hp.scalars.cpu().numpy()
    for i, w in enumerate(want['minibatches']):
        scale = max(abs(float(w[name])) for name in ('loss', 'pg', 'vl', 'entropy'))
        for j, name in enumerate(('loss', 'pg', 'vl', 'entropy')):
            assert abs(sc[i, j] - w[name]) <= REL * scale, (i, name, sc[i, j], w[name])
    # the default scan (`auto`: T split over warps at this shape, carries re-associated) stays within 1e-5 of max|returns|
    auto = ops.gae_returns(hp.rewards, hp.values, hp.last_values, hp.dones, 0.99, 0.95).cpu().numpy()
    assert np.abs(auto - want['returns']).max() <= REL * np.abs(want['returns']).max()
    flat_obs = oracle.concat_step_batches(ro.obs)[0]                 # the reference's env-major flatten copy (base.py:559-564)
    for i in sampled:
        assert np.array_equal(kept[i].cpu().numpy(), flat_obs[want['minibatches'][i]['idx']])


def _replay_agent(cls, T=16, E=8, A=6, shape=(8, 8, 4), seed=5, **kw):
    """An agent over replayed environments and a tiny torch network (both deterministic)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('make_golden', os.path.join(ROOT, 'tests', 'golden', 'make_golden.py'))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    from xagents_b200.agents import TorchModel
    rng = np.random.default_rng(seed)
    obs, rewards, dones, resets = mg._streams(rng, 4 * T, E, shape, True, 0.1)
    envs = [mg.ReplayEnv(obs[i], rewards[i], dones[i], resets[i], mg.Discrete(A)) for i in range(E)]

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.body = torch.nn.Linear(int(np.prod(shape)), 32)
            self.actor, self.critic = torch.nn.Linear(32, A), torch.nn.Linear(32, 1)

        def forward(self, x):
            h = torch.tanh(self.body(x.flatten(1)))
            return self.actor(h), self.critic(h)

    torch.manual_seed(seed)
    net = TorchModel(Tiny().cuda())
    return cls(envs, net, n_steps=T, quiet=True, seed=3, **kw), net


def test_ppo_pipeline_equals_the_literal_minibatch_loop():
    """run_ppo_epochs on the prepared pipeline (PPOHotPath in place over the agent's buffers) and the reference's literal loop
    (a subclass overriding update_gradients falls back to it) train the same network to the same weights."""
    from xagents_b200.agents import PPO

    class Literal(PPO):
        def update_gradients(self, *args):
            return super().update_gradients(*args)

    results = []
    for cls in (PPO, Literal):
        agent, net = _replay_agent(cls, mini_batches=4, ppo_epochs=2)
        perms = [np.random.default_rng(k).permutation(agent.batch_size).astype(np.int32) for k in range(2)]
        agent.permutation_source = lambda epoch: perms[epoch]
        assert agent._pipeline_applies(*[None] * 5) is False
        agent.train_step()
        torch.cuda.synchronize()
        assert (agent._pipeline is not None) == (cls is PPO)
        results.append((net.flat_param.clone(), torch.stack(list(agent.loss_history)).cpu().numpy(), agent.ro_returns.clone()))
    (p0, l0, r0), (p1, l1, r1) = results
    assert torch.equal(r0, r1)
    assert np.abs(l0 - l1).max() <= REL * np.abs(l1).max()
    assert float((p0 - p1).abs().max()) <= 2e-5


def test_train_step_fed_from_host_rollouts():
    """feeds.HostRolloutFeed as rollout_source: PPO.train_step() and A2C.train_step() on rollouts uploaded from pinned host
    buffers give what the oracle gives for those rollouts; agent.steps advances by n_steps * n_envs per step."""
    from xagents_b200 import feeds
    from xagents_b200.agents import A2C, PPO

    class Table:                                    # model outputs are inputs (precomputed per sample, env-major ids)
        output_is_softmax, comm = False, None

        def __init__(self):
            self.calls = []

        def forward(self, states, training=True):
            raise AssertionError('fed rollouts need no rollout-time forward')

        def backward_and_step(self, d_actor, d_values, grad_norm=None):
            self.calls.append((d_actor.clone(), d_values.clone()))

    T, E, A, K, M = 12, 6, 5, 2, 3
    rollouts = [synthetic.make_rollout(T, E, obs_shape=(84, 84, 4), n_actions=A, epochs=K, p_done=0.1, seed=100 + k) for k in range(3)]
    host = []
    for ro in rollouts:
        d = {k: torch.as_tensor(getattr(ro, k)).pin_memory() for k in ('obs', 'rewards', 'values', 'dones', 'actions', 'log_probs', 'last_values')}
        host.append(d)
    envs = feeds.FedEnvs(E, (84, 84, 4), torch.uint8, A, device=DEV)
    model = Table()
    agent = PPO(envs, model, n_steps=T, ppo_epochs=K, mini_batches=M, quiet=True)
    feed = feeds.HostRolloutFeed(agent, lambda k: host[k] if k < len(host) else None)
    agent.rollout_source = feed
    state = {'k': 0}
    agent.permutation_source = lambda epoch: rollouts[state['k']].permutations[epoch]

    def forward_into(states, actor_dst, critic_dst, i):
        ro = rollouts[state['k']]
        idx = ro.permutations[i // M][(i % M) * (T * E // M):(i % M + 1) * (T * E // M)]
        got = states.cpu().numpy()
        assert np.array_equal(got, oracle.concat_step_batches(ro.obs)[0][idx])       # the staged frames are this minibatch's
        actor_dst.copy_(torch.as_tensor(ro.new_logits[idx]))
        critic_dst.copy_(torch.as_tensor(ro.new_values[idx]))
    model.forward_into = forward_into
    for k, ro in enumerate(rollouts):
        state['k'] = k
        agent.train_step()
        torch.cuda.synchronize()
        assert agent.steps == (k + 1) * T * E
        want = oracle.ppo_train_step(ro.obs, ro.rewards, ro.dones, ro.values, ro.last_values, ro.actions, ro.log_probs,
                                     ro.permutations, ro.new_logits, ro.new_values, mini_batches=M)
        assert np.array_equal(agent.ro_returns.cpu().numpy(), want['returns'])
        for sc, w in zip(agent.loss_history, want['minibatches']):
            scale = max(abs(float(w[name])) for name in ('loss', 'pg', 'vl', 'entropy'))
            assert abs(float(sc[0]) - w['loss']) <= REL * scale
    assert len(model.calls) == 3 * K * M
    # A2C on the same feed mechanism
    ro = rollouts[0]
    model2 = Table()
    a2c = A2C(feeds.FedEnvs(E, (84, 84, 4), torch.uint8, A, device=DEV), model2, n_steps=T, quiet=True)
    a2c.rollout_source = feeds.HostRolloutFeed(a2c, lambda k: host[0])
    tm_logits = np.ascontiguousarray(ro.new_logits.reshape(E, T, A).swapaxes(0, 1).reshape(T * E, A))     # env-major -> time-major
    tm_values = np.ascontiguousarray(ro.new_values.reshape(E, T).T.reshape(-1))
    model2.forward_into = lambda s, a, c, i: (a.copy_(torch.as_tensor(tm_logits)), c.copy_(torch.as_tensor(tm_values)))
    a2c.train_step()
    torch.cuda.synchronize()
    ret = oracle.nstep_returns(ro.rewards, ro.dones, ro.last_values, 0.99)
    assert np.abs(a2c.ro_returns.cpu().numpy() - ret).max() <= REL * np.abs(ret).max()
    logp, ent, _ = oracle.categorical_logp_entropy(tm_logits, ro.actions.reshape(-1))
    want = oracle.a2c_loss(logp, tm_values, ent, ro.values.reshape(-1), ret.reshape(-1), 0.01, 0.5)
    assert abs(float(a2c.loss_scalars[0]) - float(want['loss'])) <= REL * abs(float(want['loss']))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs of one box')
def test_ops_and_pipeline_on_a_device_that_is_not_current():
    """Tensors on cuda:1 while the process's current device stays cuda:0: launches follow the tensors' device."""
    from xagents_b200.hotpath import PPOHotPath
    assert torch.cuda.current_device() == 0
    dev = 'cuda:1'
    T, E = 16, 8
    ro = synthetic.make_rollout(T, E, obs_shape=(84, 84, 4), epochs=1, p_done=0.05)
    d = {k: torch.as_tensor(getattr(ro, k)).to(dev) for k in ('obs', 'rewards', 'values', 'last_values', 'dones')}
    ret = ops.gae_returns(d['rewards'], d['values'], d['last_values'], d['dones'], 0.99, 0.95, mode='sequential')
    want = oracle.gae_returns(ro.rewards, ro.dones, ro.values, ro.last_values, 0.99, 0.95)
    assert ret.device == torch.device(dev) and np.array_equal(ret.cpu().numpy(), want)
    idx = torch.as_tensor(ro.permutations[0]).to(dev)
    rows = ops.gather_rows(d['obs'], idx, time_major=(T, E), mode='bulk')
    assert np.array_equal(rows.cpu().numpy(), oracle.concat_step_batches(ro.obs)[0][ro.permutations[0]])
    hp = PPOHotPath(T, E, (84, 84, 4), ro.n_actions, ppo_epochs=1, mini_batches=4, device=dev)
    hp.load(ro)
    hp.perms.copy_(torch.as_tensor(np.stack(ro.permutations)))
    hp.actor_out.normal_()
    hp.critic_out.normal_()
    hp.run()
    torch.cuda.synchronize(dev)
    assert torch.cuda.current_device() == 0 and np.array_equal(hp.returns.cpu().numpy(), want)
    assert torch.isfinite(hp.scalars).all()


def test_zero_probability_rows_stay_finite():
    """Categorical(probs=) with a saturated softmax row (a probability of exactly 0): tfp's entropy uses multiply_no_nan;
    the loss scalars, d_actor and the rollout sampler's entropy stay finite and agree with the oracle."""
    n, A = 64, 4
    rng = np.random.default_rng(3)
    probs = rng.dirichlet(np.ones(A), n).astype(np.float32)
    probs[::4] = np.eye(A, dtype=np.float32)[rng.integers(0, A, len(probs[::4]))]          # one-hot rows
    actions = probs.argmax(-1).astype(np.float32)                                           # the taken action has p > 0
    values, old_values, returns = (rng.standard_normal(n).astype(np.float32) for _ in range(3))
    old_logp = np.log(probs[np.arange(n), actions.astype(int)]) - 0.05
    adv = rng.standard_normal(n).astype(np.float32)
    cu = lambda x: torch.as_tensor(x).to(DEV)
    sc, d_actor, d_values, _ = ops.ppo_loss(cu(probs), cu(values), cu(actions), cu(old_logp), cu(old_values), cu(returns),
                                            advantages=cu(adv), actor_kind='probs')
    logp, ent, _ = oracle.categorical_logp_entropy(probs, actions, True)
    want = oracle.ppo_loss(logp, values, ent, old_values, returns, old_logp, adv, 0.1, 0.01, 0.5)
    assert torch.isfinite(sc).all() and torch.isfinite(d_actor).all() and np.isfinite(ent).all()
    scale = max(abs(float(want[k])) for k in ('loss', 'pg', 'vl', 'entropy'))
    assert abs(float(sc[0]) - want['loss']) <= REL * scale and abs(float(sc[3]) - want['entropy']) <= REL * scale
    _, _, ent_dev = ops.policy_step(cu(probs), actor_kind='probs', seed=1)
    assert np.abs(ent_dev.cpu().numpy() - ent).max() <= 1e-5


def test_debug_index_check_returns_einval_instead_of_faulting():
    """XA_DEBUG_INDICES=1 (read once per process): an out-of-range index is reported as XA_EINVAL before the TMA gather runs."""
    code = (
        "import torch, sys; sys.path.insert(0, %r)\n"
        "from xagents_b200 import ops, _ffi\n"
        "src = torch.zeros((16, 4096), dtype=torch.uint8, device='cuda')\n"
        "ok = ops.gather_rows(src, torch.tensor([0, 15, 3], dtype=torch.int32, device='cuda'), mode='bulk')\n"
        "try:\n"
        "    ops.gather_rows(src, torch.tensor([0, 16, 3], dtype=torch.int32, device='cuda'), mode='bulk')\n"
        "except _ffi.XAError as e:\n"
        "    print('XAERR', e.code, e.message)\n"
        "torch.cuda.synchronize(); print('ALIVE', int(torch.ones(1, device='cuda').item()))\n" % ROOT)
    done = subprocess.run([sys.executable, '-c', code], env=dict(os.environ, XA_DEBUG_INDICES='1'), capture_output=True, text=True, timeout=120)
    assert done.returncode == 0, done.stderr[-2000:]
    assert 'XAERR -1' in done.stdout and 'outside [0, 16)' in done.stdout and 'ALIVE 1' in done.stdout


@pytest.mark.timeout(300)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs of one box')
def test_fused_peer_allreduce_adam_two_ranks():
    """csrc/peer_adam.cu under torchrun x2: equals the oracle's clip_by_global_norm + Adam on the rank-averaged gradient,
    weights bit-identical on both ranks, no wait timed out."""
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), os.path.join(ROOT, 'scripts', 'peer_adam_check.py')]
    done = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=280)
    assert done.returncode == 0, done.stderr[-3000:]
    out = json.loads([ln for ln in done.stdout.splitlines() if ln.startswith('{')][-1])
    assert out['transport'] in ('symm', 'ipc'), out
    assert out['parity_ok'] and out['weights_identical_across_ranks'] and out['status'] == 0 and out['status_after_timing'] == 0, out
    assert out['peer_allgather_ok'], out


@pytest.mark.timeout(300)
@pytest.mark.parametrize('env_id,network', [('SyntheticAtariDevice-v0', 'torch'), ('SyntheticAtariDevice-v0', 'tcgen05'), ('CartPoleDevice-v1', 'mlp')])
def test_rollout_as_one_cuda_graph_equals_the_eager_loop(env_id, network):
    """graph_rollout: the n_steps x (policy evaluation, in-kernel sampler, device environment step, row writes) captured once
    and replayed -- same environment generator stream (registered with the graph), same Philox offsets (device counter), so the
    rollout buffers, the episode bookkeeping and agent.steps equal the eager loop's, rollout after rollout."""
    from xagents_b200 import envs
    from xagents_b200.agents import PPO, NatureCNN, TorchModel
    T, E = 6, 8
    agents = []
    for graph in (True, False):
        made = envs.create_envs(env_id, E, preprocess=env_id.startswith('Synthetic'), device=DEV)
        made.seed(5)
        if hasattr(made, 'p_done'):
            made.p_done = 0.2
        made.reset_all()
        torch.manual_seed(0)
        if network == 'mlp':
            class Mlp(torch.nn.Module):
                def __init__(self):
                    super().__init__()
                    self.body, self.actor, self.critic = torch.nn.Linear(4, 16), torch.nn.Linear(16, 2), torch.nn.Linear(16, 1)

                def forward(self, x):
                    h = torch.tanh(self.body(x))
                    return self.actor(h), self.critic(h)
            net = TorchModel(Mlp().cuda())
        else:
            net = TorchModel(NatureCNN(4, 6).cuda(), tensor_core_inference=(network == 'tcgen05'))
        agent = PPO(made, net, n_steps=T, mini_batches=4, ppo_epochs=1, quiet=True, seed=3)
        agent.graph_rollout = graph
        agents.append(agent)
    g, e = agents
    for rollout in range(3):
        for a in (g, e):
            a.get_batch()
            a._flush_episode_log()
        torch.cuda.synchronize()
        assert (g._rollout_graph is not None and g._rollout_graph is not False) and not e._rollout_graph
        for name in ('ro_states', 'ro_actions', 'ro_rewards', 'ro_dones', 'ro_values', 'ro_log_probs', 'ro_entropies', 'ro_actor'):
            assert torch.equal(getattr(g, name), getattr(e, name)), (rollout, name)
        assert torch.equal(g.get_states(), e.get_states()) and torch.equal(g.get_dones(), e.get_dones())
        assert g.steps == e.steps == (rollout + 1) * T * E and g.games == e.games and list(g.total_rewards) == list(e.total_rewards)
    assert g.games > 0 or env_id.startswith('CartPole')
    g.train_step()                                                 # and a whole train step on top of a replayed rollout
    torch.cuda.synchronize()
    assert g.steps == 4 * T * E and torch.isfinite(g.net.flat_param).all()


@pytest.mark.parametrize('case', ['grad_ppo_logits', 'grad_ppo_probs', 'grad_ppo_normal', 'grad_a2c_logits', 'grad_a2c_probs', 'grad_a2c_normal'])
def test_loss_kernel_gradients_vs_differences_of_the_reference_loss(golden, case):
    """d loss / d(actor_output, critic_output) returned by xa_ppo_loss_f32 / xa_a2c_loss_f32 against central differences of the
    loss the REFERENCE'S OWN PPO.update_gradients / A2C.train_step computes (float64 shim; tests/golden/make_golden.py
    --gradients-only) -- the three distribution branches of a2c/agent.py:50-63."""
    g = golden(case)
    kind, actor_kind = str(g['kind']), str(g['actor_kind'])
    cu = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32).to(DEV)
    common = dict(entropy_coef=float(g['entropy_coef']), value_loss_coef=float(g['value_loss_coef']), actor_kind=actor_kind)
    if kind == 'ppo':
        sc, d_actor, d_values, _ = ops.ppo_loss(cu(g['actor_output']), cu(g['critic_output']), cu(g['actions']), cu(g['old_log_probs']),
                                                cu(g['old_values']), cu(g['returns']), advantages=cu(g['advantages']),
                                                clip_norm=float(g['clip_norm']), **common)
    else:
        sc, d_actor, d_values = ops.a2c_loss(cu(g['actor_output']), cu(g['critic_output']), cu(g['actions']), cu(g['old_values']),
                                             cu(g['returns']), **common)
    assert abs(float(sc[0]) - float(g['loss'])) <= REL * max(abs(float(g['loss'])), 1.0)
    assert np.abs(d_actor.cpu().numpy() - g['d_actor']).max() <= REL * np.abs(g['d_actor']).max()
    assert np.abs(d_values.cpu().numpy() - g['d_critic']).max() <= REL * np.abs(g['d_critic']).max()


def test_native_synthetic_environment_step():
    """csrc/synth_env.cu (one launch per step of envs.BatchedSyntheticAtari on the GPU): the step_envs contract of
    xagents/base.py:408-426 -- the returned frame is the terminal one, the held frame the post-reset one, episode sums cut at
    dones --, every frame a row of the pool, rewards / dones at the configured rates, draws a function of (seed, counter + offset)."""
    from xagents_b200 import envs
    E = 512
    made = envs.create_envs('SyntheticAtariDevice-v0', E, preprocess=True, device=DEV)
    made.seed(9)
    made.p_done, made.p_reward = 0.25, 0.5
    made.reset_all()
    assert made.native_step
    pool = made._pool.flatten(1)
    sums, log = torch.zeros(E, device=DEV), torch.empty(E, device=DEV)
    want_sums = np.zeros(E)
    n_done = n_reward = n_plus = 0
    for step in range(6):
        new_states, rewards, dones = (torch.empty_like(made.states), torch.empty(E, device=DEV), torch.empty(E, device=DEV))
        made.step_into(new_states, rewards, dones, sums, log)
        r, d = rewards.cpu().numpy(), dones.cpu().numpy()
        assert set(np.unique(r)) <= {-1.0, 0.0, 1.0} and set(np.unique(d)) <= {0.0, 1.0}
        moved_on = (made.states != new_states).flatten(1).any(1)
        assert not bool((moved_on & ~dones.bool()).any())          # finished environments already hold another frame ...
        assert int((dones.bool() & ~moved_on).sum()) <= 6          # ... (the same pool row again w.p. 1/256)
        for frames in (new_states, made.states):                   # every frame is a row of the pool
            rows = frames.flatten(1)[::37]
            assert bool((rows[:, None, :64] == pool[None, :, :64]).all(-1).any(-1).all())
        want_sums += r
        assert np.array_equal(log.cpu().numpy(), want_sums.astype(np.float32))
        want_sums[d > 0] = 0
        assert np.array_equal(sums.cpu().numpy(), want_sums.astype(np.float32))
        n_done, n_reward, n_plus = n_done + int(d.sum()), n_reward + int((r != 0).sum()), n_plus + int((r > 0).sum())
    n = 6 * E
    assert abs(n_done / n - 0.25) < 0.04 and abs(n_reward / n - 0.5) < 0.04 and abs(n_plus / max(n_reward, 1) - 0.5) < 0.06
    # same (seed, counter + offset) -> same step; the device counter and the host offset add up
    outs = []
    for counter, offset in ((0, 5), (5, 0), (2, 3), (0, 6)):
        made._counter.fill_(counter)
        made._host_offset = offset
        new_states, rewards, dones = made.step_all(None)
        outs.append((new_states.clone(), rewards.clone(), dones.clone(), made.states.clone()))
    for other in outs[1:3]:
        assert all(torch.equal(a, b) for a, b in zip(outs[0], other))
    assert not torch.equal(outs[0][0], outs[3][0])
    assert ops.launch_count() >= 0
    with pytest.raises(Exception, match='different buffers'):
        made.step_into(made.states, rewards, dones)


@pytest.mark.timeout(300)
@pytest.mark.parametrize('network', ['torch', 'tcgen05'])
def test_fused_rollout_steps_equal_the_generic_loop(network):
    """A2C._rollout_loop_fused (network, sampler and environment kernel write rows t / t+1 of the rollout buffers in place: three
    native calls per step) against the generic loop of a2c/agent.py:113-139 on the same environment and seeds: identical buffers,
    states, dones and episode bookkeeping, rollout after rollout."""
    from xagents_b200 import envs
    from xagents_b200.agents import PPO, NatureCNN, TorchModel
    T, E = 5, 16
    agents = []
    for fused in (True, False):
        made = envs.create_envs('SyntheticAtariDevice-v0', E, preprocess=True, device=DEV)
        made.seed(5)
        made.p_done = 0.2
        made.reset_all()
        torch.manual_seed(0)
        net = TorchModel(NatureCNN(4, 6).cuda(), tensor_core_inference=(network == 'tcgen05'))
        agent = PPO(made, net, n_steps=T, mini_batches=4, ppo_epochs=1, quiet=True, seed=3)
        agent.graph_rollout = False
        if not fused:
            agent._fused_rollout_applies = lambda: False
        agents.append(agent)
    f, g = agents
    assert f._fused_rollout_applies()
    for rollout in range(3):
        for a in (f, g):
            a.get_batch()
            a._flush_episode_log()
        torch.cuda.synchronize()
        for name in ('ro_states', 'ro_actions', 'ro_rewards', 'ro_dones', 'ro_values', 'ro_log_probs', 'ro_entropies', 'ro_actor', 'ro_returns'):
            assert torch.equal(getattr(f, name), getattr(g, name)), (rollout, name)
        assert torch.equal(f.get_states(), g.get_states()) and torch.equal(f.get_dones(), g.get_dones())
        assert f.steps == g.steps == (rollout + 1) * T * E and f.games == g.games and list(f.total_rewards) == list(g.total_rewards)
    assert f.games > 0


@pytest.mark.timeout(300)
def test_ppo_without_a_frame_gather_equals_ppo_with_it():
    """PPO.run_ppo_epochs with a network whose first layer fetches every frame through the permutation
    (`TorchModel.reads_through_permutation`: PPOHotPath(obs_gather=False), no staged copy of the minibatches) against the
    same agent with the gather launch + the same network on the staged rows: identical weights and loss scalars."""
    from xagents_b200 import feeds
    from xagents_b200.agents import PPO, NatureCnnTc, TorchModel
    T, E, A = 16, 24, 6
    results = []
    for gather in (False, True):
        torch.manual_seed(11)
        envs = feeds.FedEnvs(E, (84, 84, 4), torch.uint8, A, device=DEV)
        net = TorchModel(NatureCnnTc(4, A).cuda())
        assert net.reads_through_permutation
        agent = PPO(envs, net, n_steps=T, mini_batches=3, ppo_epochs=2, quiet=True, seed=5, device=DEV)
        agent.pipeline_options = dict(obs_gather=gather)
        gen = torch.Generator(device=DEV)
        gen.manual_seed(2)
        agent.ro_states.copy_(torch.randint(0, 256, agent.ro_states.shape, dtype=torch.uint8, device=DEV, generator=gen))
        for name, scale in (('ro_rewards', 1.0), ('ro_values', 0.5), ('ro_log_probs', 0.3)):
            getattr(agent, name).copy_(torch.randn(getattr(agent, name).shape, device=DEV, generator=gen) * scale)
        agent.ro_log_probs.abs_().neg_()
        agent.ro_dones.copy_((torch.rand(agent.ro_dones.shape, device=DEV, generator=gen) < 0.1).float())
        agent.ro_actions.copy_(torch.randint(0, A, agent.ro_actions.shape, device=DEV, generator=gen).float())
        agent.ro_returns.copy_(torch.randn(agent.ro_returns.shape, device=DEV, generator=gen))
        batch = agent.concat_step_batches(agent.ro_states, agent.ro_actions, agent.ro_returns, agent.ro_values, agent.ro_log_probs)
        for _ in range(2):
            agent.run_ppo_epochs(*batch)
        torch.cuda.synchronize()
        hp = agent.hot_path()
        assert hp.obs_gather == gather and (hp.mb_obs is None) == (not gather)
        results.append((net.flat_param.clone(), torch.stack(list(agent.loss_history)).clone(), net.step))
    (p0, l0, s0), (p1, l1, s1) = results
    assert s0 == s1 == 2 * 2 * 3 and torch.isfinite(p0).all()
    assert torch.equal(l0, l1) and torch.equal(p0, p1)


@pytest.mark.timeout(300)
def test_chained_launches_do_not_change_results():
    """Programmatic dependent launches (csrc/xa_common.cuh `launch_chained`; xa_set_chained_launches): a kernel of the rollout
    step / network pass / loss -> optimiser chain may be scheduled while its predecessor drains, and waits for the predecessor's
    completion before its first global access.  Rollouts (fused three-call steps on the tcgen05 network) followed by PPO update
    phases must give bit-identical buffers, weights and loss scalars at level 0 (plain stream order), 1 (default policy) and 2
    (every launch chained)."""
    from xagents_b200 import _ffi, envs
    from xagents_b200.agents import PPO, NatureCnnTc, TorchModel
    lib = _ffi.lib()
    T, E, A = 8, 32, 6
    before = lib.xa_set_chained_launches(-1)
    assert before in (0, 1, 2, 3)
    results = []
    try:
        for level in (0, 1, 2):
            assert lib.xa_set_chained_launches(level) in (0, 1, 2, 3) and lib.xa_set_chained_launches(-1) == level
            made = envs.create_envs('SyntheticAtariDevice-v0', E, preprocess=True, device=DEV)
            made.seed(5)
            made.p_done = 0.1
            made.reset_all()
            torch.manual_seed(0)
            net = TorchModel(NatureCnnTc(4, A).cuda())
            agent = PPO(made, net, n_steps=T, mini_batches=4, ppo_epochs=2, quiet=True, seed=3)
            agent.graph_rollout = False
            for _ in range(3):
                agent.train_step()
            torch.cuda.synchronize()
            results.append([t.clone() for t in (agent.ro_states, agent.ro_actions, agent.ro_values, agent.ro_log_probs, agent.ro_returns, net.flat_param,
                                                torch.stack(list(agent.loss_history)))])
    finally:
        lib.xa_set_chained_launches(before)
    assert torch.isfinite(results[0][5]).all()
    for other in results[1:]:
        for a, b in zip(results[0], other):
            assert torch.equal(a, b)
