"""The drop-in PPO / A2C classes against fixtures produced by the reference's own PPO / A2C.train_step
(tests/golden/make_golden.py): same replayed environments, same model weights, same sampled actions and
the same shuffles, so every intermediate of the train step can be compared."""
import importlib.util
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
REL = 1e-5
HERE = os.path.dirname(os.path.abspath(__file__))


def _golden_module():
    spec = importlib.util.spec_from_file_location('make_golden', os.path.join(HERE, 'golden', 'make_golden.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)                      # defines helpers only; the reference is NOT imported
    return mod


class NumpyModel:
    """Adapter around the generator's TinyModel so model outputs are bit-identical to the fixture run."""
    output_is_softmax = False
    comm = None

    def __init__(self, tiny, image):
        self.tiny, self.image, self.grads = tiny, image, []

    def forward(self, states, training=True):
        x = states.detach().cpu().numpy().astype(np.float32)
        if self.image:
            x = x / np.float32(255.0)
        actor, critic = self.tiny(x)
        return (torch.as_tensor(np.ascontiguousarray(actor)).cuda(),
                torch.as_tensor(np.ascontiguousarray(critic.reshape(-1))).cuda())

    def backward_and_step(self, d_actor, d_values, grad_norm=None):
        self.grads.append((d_actor.clone(), d_values.clone()))


def _softmax_tiny(mg, in_features, n_actions, rng):
    """TinyModel(softmax=True) outside the generator: its constructor looks the Keras softmax symbol up in the stubbed
    tensorflow module, which only exists while make_golden.py runs; the arithmetic needs the flag alone."""
    import sys
    import types
    had = 'tensorflow' in sys.modules
    if not had:
        sys.modules['tensorflow'] = types.SimpleNamespace(keras=types.SimpleNamespace(activations=types.SimpleNamespace(softmax='softmax')))
    try:
        return mg.TinyModel(in_features, n_actions, rng, softmax=True)
    finally:
        if not had:
            del sys.modules['tensorflow']


CASES = {
    'ppo_image': dict(seed=11, T=16, E=8, shape=(8, 8, 4), image=True, A=6, p=0.08, kw=dict(mini_batches=4, ppo_epochs=4), drift=0.05),
    'ppo_cartpole': dict(seed=12, T=128, E=16, shape=(4,), image=False, A=2, p=0.02, kw=dict(mini_batches=4, ppo_epochs=4), drift=0.05),
    'ppo_ragged': dict(seed=13, T=7, E=3, shape=(5,), image=False, A=3, p=0.3, kw=dict(mini_batches=4, ppo_epochs=2, clip_norm=0.02), drift=0.2),
    'ppo_single_env': dict(seed=14, T=9, E=1, shape=(6, 6, 1), image=True, A=4, p=0.2, kw=dict(mini_batches=3, ppo_epochs=2), drift=0.05),
    'ppo_box': dict(seed=15, T=10, E=4, shape=(6,), image=False, A=3, p=0.15, kw=dict(mini_batches=4, ppo_epochs=2), drift=0.05, box=True),
    'ppo_softmax': dict(seed=16, T=12, E=6, shape=(5,), image=False, A=5, p=0.1, kw=dict(mini_batches=3, ppo_epochs=2), drift=0.05, softmax=True),
    'a2c_box': dict(seed=23, T=6, E=5, shape=(4,), image=False, A=2, p=0.2, kw={}, drift=0.05, box=True),
    'a2c_softmax': dict(seed=24, T=7, E=4, shape=(4,), image=False, A=3, p=0.2, kw={}, drift=0.05, softmax=True),
    'a2c_image': dict(seed=21, T=5, E=16, shape=(8, 8, 4), image=True, A=6, p=0.1, kw={}, drift=0.05),
    'a2c_vector': dict(seed=22, T=12, E=5, shape=(4,), image=False, A=2, p=0.15, kw={}, drift=0.05),
}


def build(case, cls):
    mg = _golden_module()
    c = CASES[case]
    rng = np.random.default_rng(c['seed'])
    obs, rewards, dones, resets = mg._streams(rng, c['T'], c['E'], c['shape'], c['image'], c['p'])
    space = mg.Box((c['A'],)) if c.get('box') else mg.Discrete(c['A'])
    envs = [mg.ReplayEnv(obs[i], rewards[i], dones[i], resets[i], space) for i in range(c['E'])]
    in_features = min(int(np.prod(c['shape'])), 24)
    tiny = _softmax_tiny(mg, in_features, c['A'], rng) if c.get('softmax') else mg.TinyModel(in_features, c['A'], rng)
    model = NumpyModel(tiny, c['image'])
    model.output_is_softmax = bool(c.get('softmax'))
    agent = cls(envs, model, n_steps=c['T'], quiet=True, **c['kw'])
    return agent, model, tiny, c


def near(got, want, scale=None):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    want = np.asarray(want, np.float64).reshape(got.shape)
    s = np.abs(want).max() if scale is None else scale
    assert np.abs(got - want).max() <= REL * max(s, 1e-30)


@pytest.mark.parametrize('case', ['ppo_image', 'ppo_cartpole', 'ppo_ragged', 'ppo_single_env', 'ppo_box', 'ppo_softmax'])
def test_ppo_train_step_matches_the_reference_run(golden, case):
    from xagents_b200.agents import PPO
    g = golden(case)
    agent, model, tiny, c = build(case, PPO)
    T, E = c['T'], c['E']
    tm = lambda flat: np.ascontiguousarray(np.asarray(flat).reshape(E, T).T)
    golden_actions = np.ascontiguousarray(np.swapaxes(g['flat_actions'].reshape((E, T) + ((c['A'],) if c.get('box') else ())), 0, 1))
    agent.action_source = lambda step, actor_out: golden_actions[step]
    assert agent.actor_kind == ('normal' if c.get('box') else 'probs' if c.get('softmax') else 'logits')
    agent.permutation_source = lambda epoch: g['shuffles'][epoch]
    inner = agent.run_ppo_epochs

    def run_epochs(*batch):                           # weights "move" between rollout and updates, as in the fixture
        tiny.drift = np.float32(c['drift'])
        return inner(*batch)

    agent.run_ppo_epochs = run_epochs
    assert agent.batch_size == T * E and agent.mini_batch_size == int(g['mini_batch_size'])
    agent.train_step()
    torch.cuda.synchronize()
    assert agent.steps == int(g['steps_after']) == T * E
    # rollout assembly (a1): time-major buffers hold what the reference's lists held
    assert np.array_equal(agent.ro_states.cpu().numpy().astype(np.float32), g['states_time_major'])
    assert np.array_equal(agent.ro_rewards.cpu().numpy(), g['rewards'])
    assert np.array_equal(agent.ro_dones.cpu().numpy(), g['dones'])
    assert np.array_equal(agent.ro_values.cpu().numpy(), g['values'].reshape(T, E))
    near(agent.ro_log_probs, tm(g['flat_log_probs']))
    near(agent.ro_returns, g['returns'])
    # every update of every epoch
    assert len(agent.loss_history) == len(g['losses']) == len(model.grads)
    for i, sc in enumerate(agent.loss_history):
        sc = sc.cpu().numpy()
        _, ref_ent, ref_vl, ref_pg = g['means'][i]
        scale = max(abs(float(g['losses'][i])), abs(ref_ent), abs(0.5 * ref_vl), abs(ref_pg))
        assert abs(sc[0] - g['losses'][i]) <= REL * scale, (i, sc, g['losses'][i])
        assert abs(sc[1] - ref_pg) <= REL * scale and abs(sc[2] - 0.5 * ref_vl) <= REL * scale and abs(sc[3] - ref_ent) <= REL * scale


def test_ppo_get_mini_batches_contract(golden):
    """list(get_mini_batches(...)) is the reference's list: K*ceil(N/B) minibatches of 5 gathered items."""
    from xagents_b200.agents import PPO
    g = golden('ppo_ragged')
    agent, model, tiny, c = build('ppo_ragged', PPO)
    T, E = c['T'], c['E']
    golden_actions = np.ascontiguousarray(g['flat_actions'].reshape(E, T).T)
    agent.action_source = lambda step, actor_out: golden_actions[step]
    agent.permutation_source = lambda epoch: g['shuffles'][epoch]
    batch = agent.get_batch()
    assert [tuple(b.shape) for b in batch] == [(T * E,) + c['shape'], (T * E,), (T * E,), (T * E,), (T * E,)]
    assert np.array_equal(batch[0].materialize().cpu().numpy(), g['flat_states'])
    mbs = list(agent.get_mini_batches(*batch))
    assert len(mbs) == len(g['losses']) == 10
    for i, mb in enumerate(mbs):
        assert np.array_equal(mb[0].cpu().numpy(), g[f'mb{i}_states'])
        assert np.array_equal(mb[1].cpu().numpy(), g[f'mb{i}_actions'].reshape(-1))
        assert np.array_equal(mb[3].cpu().numpy(), g[f'mb{i}_old_values'].reshape(-1))
    assert len(mbs[4][0]) == 1                        # trailing short minibatch: 21 = 4*5 + 1


@pytest.mark.parametrize('case', ['a2c_image', 'a2c_vector', 'a2c_box', 'a2c_softmax'])
def test_a2c_train_step_matches_the_reference_run(golden, case):
    from xagents_b200.agents import A2C
    g = golden(case)
    agent, model, tiny, c = build(case, A2C)
    T, E = c['T'], c['E']
    golden_actions = np.ascontiguousarray(np.swapaxes(g['flat_actions'].reshape((E, T) + ((c['A'],) if c.get('box') else ())), 0, 1))
    assert agent.actor_kind == ('normal' if c.get('box') else 'probs' if c.get('softmax') else 'logits')
    agent.action_source = lambda step, actor_out: golden_actions[step]
    inner = agent.calculate_returns

    def returns_then_drift(*a, **k):
        out = inner(*a, **k)
        tiny.drift = np.float32(c['drift'])
        return out

    agent.calculate_returns = returns_then_drift
    agent.train_step()
    torch.cuda.synchronize()
    assert agent.steps == int(g['steps_after'])
    assert np.array_equal(agent.ro_rewards.cpu().numpy(), g['rewards'])
    assert np.array_equal(agent.ro_dones.cpu().numpy(), g['dones'])
    sc = agent.loss_scalars.cpu().numpy()
    assert abs(sc[0] - g['loss'][0]) <= REL * max(abs(float(g['loss'][0])), abs(sc[2]), abs(sc[3]))


def test_constructor_contract_and_errors():
    from xagents_b200.agents import A2C, PPO
    mg = _golden_module()
    with pytest.raises(AssertionError, match='No environments given'):
        PPO([], None)
    agent, *_ = build('ppo_ragged', PPO)
    for name, want in (('lam', 0.95), ('ppo_epochs', 2), ('mini_batches', 4), ('advantage_epsilon', 1e-8), ('clip_norm', 0.02),
                       ('entropy_coef', 0.01), ('value_loss_coef', 0.5), ('grad_norm', 0.5), ('gamma', 0.99), ('n_envs', 3)):
        assert getattr(agent, name) == want
    with pytest.raises(AssertionError, match='Invalid batch size to mini-batch size ratio'):
        rng = np.random.default_rng(0)
        obs, rewards, dones, resets = mg._streams(rng, 1, 1, (4,), False, 0.0)
        PPO([mg.ReplayEnv(obs[0], rewards[0], dones[0], resets[0], mg.Discrete(2))],
            NumpyModel(mg.TinyModel(4, 2, rng), False), n_steps=1, mini_batches=4)
    with pytest.raises(AssertionError, match='should be specified when fit'):
        agent.fit()
    with pytest.raises(NotImplementedError):
        from xagents_b200.agents import BaseAgent
        BaseAgent.train_step(agent)


def test_fit_with_a_torch_model_trains_on_device():
    """fit(max_steps=...) end to end: NatureCNN-shaped torch model, fused clip+Adam, stop condition."""
    from xagents_b200.agents import PPO, NatureCNN, TorchModel
    mg = _golden_module()
    T, E, A = 8, 4, 6
    rng = np.random.default_rng(7)
    obs, rewards, dones, resets = mg._streams(rng, 4 * T, E, (84, 84, 4), True, 0.1)
    envs = [mg.ReplayEnv(obs[i], rewards[i], dones[i], resets[i], mg.Discrete(A)) for i in range(E)]
    torch.manual_seed(0)
    net = TorchModel(NatureCNN(4, A).cuda())
    before = net.flat_param.clone()
    agent = PPO(envs, net, n_steps=T, mini_batches=4, ppo_epochs=2, quiet=True, seed=3)
    agent.fit(max_steps=3 * T * E)
    torch.cuda.synchronize()
    assert agent.steps == 3 * T * E and net.step == 3 * 2 * 4
    assert torch.isfinite(net.flat_param).all() and not torch.equal(before, net.flat_param)
    assert net.n_params == 1_687_719 - 0 or net.n_params > 1_600_000      # Nature CNN @84x84x4, 6 actions
    losses = torch.stack(agent.loss_history).cpu().numpy()
    assert np.isfinite(losses).all()


@pytest.mark.timeout(240)
def test_ppo_fit_with_the_as_coded_conv1d_network_on_the_tensor_core_dense_path():
    """The .cfg as ModelReader actually builds it (Conv1D trunk, 19.3 M parameters): same agent, loss kernels and fused
    optimiser; its 37 632 -> 512 Dense layer runs on the tcgen05 GEMM."""
    import importlib.util
    import os

    import numpy as np
    from xagents_b200.agents import PPO, AsCodedConv1dCNN, TorchModel
    spec = importlib.util.spec_from_file_location('make_golden', os.path.join(os.path.dirname(__file__), 'golden', 'make_golden.py'))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    T, E, A = 8, 4, 6
    rng = np.random.default_rng(3)
    obs, rewards, dones, resets = mg._streams(rng, 3 * T, E, (84, 84, 4), True, 0.1)
    envs = [mg.ReplayEnv(obs[i], rewards[i], dones[i], resets[i], mg.Discrete(A)) for i in range(E)]
    torch.manual_seed(0)
    net = TorchModel(AsCodedConv1dCNN(4, A, tensor_core_dense=True).cuda())
    assert net.n_params == 19_293_351
    before = net.flat_param.clone()
    agent = PPO(envs, net, n_steps=T, mini_batches=4, ppo_epochs=2, quiet=True, seed=5)
    agent.fit(max_steps=2 * T * E)
    torch.cuda.synchronize()
    assert agent.steps == 2 * T * E and net.step == 2 * 2 * 4
    assert torch.isfinite(net.flat_param).all() and not torch.equal(before, net.flat_param)


@pytest.mark.timeout(300)
def test_cli_train_ppo_on_cartpole_learns():
    """BASELINE config C1 end to end through the reference's command line: `xagents train ppo --env CartPole-v1
    --n-envs 16` (n_steps defaults to 128), default `.cfg` network, every hot-path stage on the device.  A random policy
    scores ~22 per episode; the mean over the last 100 episodes must clearly exceed that."""
    from xagents_b200 import cli
    ex = cli.Executor()
    ex.execute(['train', 'ppo', '--env', 'CartPole-v1', '--n-envs', '16', '--max-steps', '81920', '--seed', '1', '--quiet'])
    agent = ex.agent
    assert type(agent).__name__ == 'PPO' and agent.n_steps == 128 and agent.n_envs == 16 and agent.steps >= 81920
    assert agent.net.n_params == 4675 and agent.net.step == (81920 // 2048) * 16
    agent.update_metrics()
    assert torch.isfinite(agent.net.flat_param).all()
    assert agent.best_reward > 35.0, f'PPO did not improve on CartPole: best mean reward {agent.best_reward}'


@pytest.mark.timeout(300)
def test_cli_train_a2c_on_synthetic_atari_frames_with_both_cfg_readings():
    """Config C2's shape through the command line (A2C, 84x84x4 uint8 frames, n_envs=16, n_steps=5), the default `.cfg`
    read as Conv2D (documented) and as Conv1D (as coded), Dense layers on the tcgen05 GEMM."""
    from xagents_b200 import cli
    for dims, n_params in ((2, 1_687_719), (1, 19_293_351)):
        ex = cli.Executor()
        ex.execute(['train', 'a2c', '--env', 'SyntheticAtari-v0', '--n-envs', '16', '--max-steps', '400', '--seed', '2', '--quiet',
                    '--conv-dims', str(dims), '--tensor-core-dense', '--preprocess'])
        agent = ex.agent
        assert agent.n_steps == 5 and agent.steps == 400 and agent.net.n_params == n_params and agent.net.step == 5
        assert agent.ro_states.dtype == torch.uint8 and tuple(agent.ro_states.shape) == (5, 16, 84, 84, 4)
        assert torch.isfinite(agent.net.flat_param).all()


@pytest.mark.timeout(300)
def test_cli_train_ppo_with_the_tensor_core_network_on_the_device_environment():
    """`--tensor-core-network`: the default cnn `.cfg` as `NatureCnnTc` (tcgen05 forward and backward, the file's initialisers and
    seed) through the command line, on the device-resident synthetic Atari environments: rollout as one CUDA graph of fused steps,
    update phase without a frame gather (the first layer reads the rollout through the permutation)."""
    from xagents_b200 import cli
    ex = cli.Executor()
    ex.execute(['train', 'ppo', '--env', 'SyntheticAtariDevice-v0', '--n-envs', '8', '--n-steps', '16', '--max-steps', '384', '--seed', '2',
                '--quiet', '--tensor-core-network', '--preprocess', '--mini-batches', '2', '--ppo-epochs', '2'])
    agent = ex.agent
    assert type(agent.net.module).__name__ == 'NatureCnnTc' and agent.net.n_params == 1_687_719 and agent.net.reads_through_permutation
    assert agent.steps == 384 and agent.net.step == 3 * 2 * 2 and torch.isfinite(agent.net.flat_param).all()
    assert agent._rollout_graph and agent._fused_rollout_applies() and agent._pipeline is not None and not agent._pipeline.obs_gather
    assert agent._pipeline.mb_obs is None


@pytest.mark.timeout(300)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs of one box')
def test_cli_train_shards_the_environments_over_two_gpus():
    """torchrun x2: 16 CartPole environments, 8 per rank, NCCL gradient all-reduce per minibatch and job-wide advantage
    moments.  Every rank must end with bit-identical weights after the same number of optimiser steps."""
    import json
    import socket
    import subprocess
    import sys
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    root = os.path.dirname(HERE)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(root, 'scripts', 'sharded_train_check.py'), 'train', 'ppo', '--env', 'CartPole-v1',
           '--n-envs', '16', '--max-steps', '40960', '--seed', '1', '--quiet']
    done = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=280)
    assert done.returncode == 0, done.stderr[-3000:]
    line = [ln for ln in done.stdout.splitlines() if ln.startswith('{')][-1]
    out = json.loads(line)
    assert out['world_size'] == 2 and out['n_envs_per_rank'] == 8 and out['job_steps'] == 40960
    assert out['optimizer_steps'] == (40960 // 2048) * 16
    assert out['weights_identical_across_ranks'] and out['weights_finite']
    assert out['mean_reward_last_episodes'] > 25.0


@pytest.mark.timeout(240)
def test_ppo_and_acer_on_the_device_resident_batched_environment():
    """envs.BatchedSyntheticAtari: the rollout never leaves the device (states, rewards, dones are tensors; no per-env loop),
    episode bookkeeping is read back once per train step and agrees with the done flags the rollout recorded."""
    from xagents_b200 import cli, envs
    from xagents_b200.agents import PPO, NatureCNN, TorchModel
    made = envs.create_envs('SyntheticAtariDevice-v0', 8, preprocess=True, device='cuda:0')
    made.p_done = 0.2
    torch.manual_seed(0)
    net = TorchModel(NatureCNN(4, 6).cuda())
    agent = PPO(made, net, n_steps=8, mini_batches=4, ppo_epochs=2, quiet=True, seed=3)
    assert agent.batched and agent.obs_dtype == torch.uint8 and agent.get_states().is_cuda
    first = agent.get_states().clone()
    agent.train_step()
    assert torch.equal(agent.ro_states[0], first) and agent.steps == 64 and agent.games == 0
    agent.check_episodes() if agent.training_start_time else agent._flush_episode_log()
    assert agent.games == int(agent.ro_dones[1:].sum().item()) > 0 and len(agent.total_rewards) == agent.games
    assert torch.equal(agent.ro_dones[8], agent.get_dones()) and set(agent.ro_rewards.unique().tolist()) <= {-1.0, 0.0, 1.0}
    agent.fit(max_steps=3 * 64)
    assert agent.steps == 192 and net.step == 3 * 2 * 4 and torch.isfinite(net.flat_param).all()
    # through the command line, with ACER's device trajectory ring fed by the device environment
    ex = cli.Executor()
    ex.execute(['train', 'acer', '--env', 'SyntheticAtariDevice-v0', '--n-envs', '4', '--n-steps', '8', '--max-steps', '96', '--quiet',
                '--seed', '3', '--buffer-max-size', '16', '--buffer-initial-size', '8', '--conv-dims', '2', '--preprocess'])
    assert ex.agent.batched and ex.agent.steps == 96 and ex.agent.ring.states.dtype == torch.uint8
    assert torch.isfinite(ex.agent.net.flat_param).all()


@pytest.mark.timeout(300)
def test_cli_train_ppo_on_the_device_resident_cartpole_learns():
    """Config C1 with the environments integrated on the device (envs.BatchedCartPole): no host round trip in the rollout."""
    from xagents_b200 import cli
    ex = cli.Executor()
    ex.execute(['train', 'ppo', '--env', 'CartPoleDevice-v1', '--n-envs', '16', '--max-steps', '81920', '--seed', '1', '--quiet'])
    agent = ex.agent
    assert agent.batched and agent.obs_dtype == torch.float32 and tuple(agent.ro_states.shape) == (128, 16, 4)
    assert agent.steps >= 81920 and agent.net.step == (81920 // 2048) * 16
    agent.update_metrics()
    assert agent.games > 100 and agent.best_reward > 35.0, f'PPO did not improve: best mean reward {agent.best_reward}'
