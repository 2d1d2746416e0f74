"""Host logic of the drop-in surface around the hot path (SURVEY.md §8b "CLI / factory" row): the `.cfg` network reader,
the built-in environments and the `train` command line.  No GPU: nothing here launches a kernel."""
import math
import os
import textwrap
import warnings

import numpy as np
import pytest
import torch

from xagents_b200 import cli, envs
from xagents_b200.agents import AsCodedConv1dCNN, ModelReader

MODELS = os.path.join(os.path.dirname(os.path.abspath(cli.__file__)), 'agents', 'models')
CNN = os.path.join(MODELS, 'cnn-actor-critic.cfg')
ANN = os.path.join(MODELS, 'ann-actor-critic.cfg')


def _cfg(tmp_path, text):
    path = tmp_path / 'net.cfg'
    path.write_text(textwrap.dedent(text))
    return str(path)


# ---------------------------------------------------------------------- .cfg reader
def test_default_cnn_cfg_has_the_parameter_counts_of_both_readings():
    """As coded (Conv1D over the rows): 19 293 351; as documented (Conv2D, README.md:243-259): 1 687 719."""
    as_coded = ModelReader(CNN, [6, 1], (84, 84, 4)).build_model()
    documented = ModelReader(CNN, [6, 1], (84, 84, 4), conv_dims=2).build_model()
    assert sum(p.numel() for p in as_coded.parameters()) == 19_293_351
    assert sum(p.numel() for p in documented.parameters()) == 1_687_719
    # README.md:243-259 prints 1 681 575: the same network on AtariWrapper's single-channel frames
    assert sum(p.numel() for p in ModelReader(CNN, [6, 1], (84, 84, 1), conv_dims=2).build_model().parameters()) == 1_681_575
    x = torch.rand(3, 84, 84, 4)
    for net in (as_coded, documented):
        actor, critic = net(x)
        assert actor.shape == (3, 6) and critic.shape == (3, 1) and not net.output_is_softmax


def test_as_coded_reading_equals_the_hand_written_conv1d_network():
    """Same weights -> same outputs as agents.AsCodedConv1dCNN (the module the GPU tests train)."""
    torch.manual_seed(0)
    net = ModelReader(CNN, [6, 1], (84, 84, 4)).build_model()
    ref = AsCodedConv1dCNN(4, 6)
    convs = [m for m in ref.trunk if hasattr(m, 'weight')]
    with torch.no_grad():
        for i, conv in enumerate(convs):                           # Conv2d (1, k) kernels <- Conv1d kernels
            conv.weight.copy_(net.layers[i].weight.unsqueeze(2))
            conv.bias.copy_(torch.randn_like(conv.bias))
            net.layers[i].bias.copy_(conv.bias)
        for mine, theirs in ((net.layers[4], ref.fc[0]), (net.layers[5], ref.actor), (net.layers[6], ref.critic)):
            theirs.weight.copy_(mine.weight)
            theirs.bias.copy_(torch.randn_like(theirs.bias))
            mine.bias.copy_(theirs.bias)
    x = torch.rand(2, 84, 84, 4)
    got, want = net(x), ref(x)
    for g, w in zip(got, want):
        assert torch.allclose(g, w, rtol=1e-5, atol=1e-6)


def test_tensor_core_network_starts_from_the_weights_the_reader_gives_the_cfg():
    """ModelReader.build_tensor_core_network (`--tensor-core-network`): the documented Conv2D network as `NatureCnnTc`, its
    parameters the ones the reader's initialisers and seed produce for the file -- the FC weight moved from Keras' (h, w, c)
    flatten order to the module's (c, h, w) -- so both modules compute the same function; other architectures are refused."""
    from xagents_b200.agents import NatureCNN
    ref = ModelReader(CNN, [5, 1], (84, 84, 4), conv_dims=2, seed=7).build_model()
    net = ModelReader(CNN, [5, 1], (84, 84, 4), conv_dims=2, seed=7).build_tensor_core_network()
    assert type(net).__name__ == 'NatureCnnTc' and net.takes_uint8 and not net.output_is_softmax
    assert sum(p.numel() for p in net.parameters()) == sum(p.numel() for p in ref.parameters())
    x = torch.rand(3, 84, 84, 4)
    with torch.no_grad():
        actor, critic = NatureCNN.forward(net, x)                  # the module's fp32 torch form (its own forward runs the kernels)
        want_actor, want_critic = ref(x)
    assert torch.allclose(actor, want_actor, atol=1e-5) and torch.allclose(critic, want_critic, atol=1e-5)
    with pytest.raises(ValueError, match='not the documented 84x84x4 Nature CNN'):
        ModelReader(CNN, [5, 1], (84, 84, 1), conv_dims=2).build_tensor_core_network()
    with pytest.raises(ValueError, match='not the documented 84x84x4 Nature CNN'):
        ModelReader(CNN, [12, 1], (84, 84, 4), conv_dims=2).build_tensor_core_network()       # more actions than the heads kernel holds


def test_conv2d_reading_matches_a_channels_last_restatement():
    net = ModelReader(CNN, [4, 1], (84, 84, 4), conv_dims=2, seed=3).build_model()
    x = torch.rand(2, 84, 84, 4)
    h = x.permute(0, 3, 1, 2)
    for layer, stride in zip(net.layers[:3], (4, 2, 1)):
        h = torch.relu(torch.nn.functional.conv2d(h, layer.weight, layer.bias, stride))
    assert h.shape == (2, 64, 7, 7)
    flat = h.permute(0, 2, 3, 1).reshape(2, -1)                    # Keras flatten order: (row, column, channel)
    trunk = torch.relu(flat @ net.layers[4].weight.T + net.layers[4].bias)
    actor, critic = net(x)
    assert torch.allclose(actor, trunk @ net.layers[5].weight.T + net.layers[5].bias, atol=1e-6)
    assert torch.allclose(critic, trunk @ net.layers[6].weight.T + net.layers[6].bias, atol=1e-6)


def test_ann_cfg_and_initialisers():
    net = ModelReader(ANN, [2, 1], (4,), seed=7).build_model()
    assert sum(p.numel() for p in net.parameters()) == 4 * 64 + 64 + 64 * 64 + 64 + 64 * 2 + 2 + 64 + 1
    again = ModelReader(ANN, [2, 1], (4,), seed=7).build_model()
    other = ModelReader(ANN, [2, 1], (4,), seed=8).build_model()
    assert all(torch.equal(a, b) for a, b in zip(net.parameters(), again.parameters()))
    assert not torch.equal(net.layers[0].weight, other.layers[0].weight)
    for layer, gain in zip(net.layers, (2 ** 0.5, 2 ** 0.5, 0.01, 1.0)):  # orthogonal(gain): W W^T or W^T W = gain^2 I
        w = layer.weight.double()
        gram = w @ w.T if w.shape[0] <= w.shape[1] else w.T @ w
        assert torch.allclose(gram, gain ** 2 * torch.eye(gram.shape[0], dtype=torch.float64), atol=1e-5)
        assert torch.count_nonzero(layer.bias) == 0
    actor, critic = net(torch.rand(5, 4))
    assert actor.shape == (5, 2) and critic.shape == (5, 1)


def test_common_layer_wiring_outputs_and_softmax_detection(tmp_path):
    """build_model (common.py:257-290): after a `common` layer every later dense layer reads it, not its predecessor."""
    path = _cfg(tmp_path, """
        [dense-0]
        units=8
        activation=tanh
        common=1
        [dense-1]
        units=5
        activation=relu
        output=1
        [dense-2]
        activation=softmax
        output=1
        [dense-3]
        output=1
        """)
    net = ModelReader(path, [3, 1], (6,)).build_model()
    assert [layer.linear.in_features for layer in net.layers] == [6, 8, 8, 8]      # dense-2/3 read dense-0, not dense-1
    outs = net(torch.rand(4, 6))
    assert [tuple(o.shape) for o in outs] == [(4, 5), (4, 3), (4, 1)]
    assert torch.allclose(outs[1].sum(-1), torch.ones(4), atol=1e-6)
    assert net.output_is_softmax                                                    # a2c/agent.py:42-43
    single = ModelReader(_cfg(tmp_path, '[dense-0]\nunits=3\noutput=1\n'), [], (6,)).build_model()
    assert single(torch.rand(2, 6)).shape == (2, 3)


def test_reader_errors(tmp_path):
    with pytest.raises(AssertionError, match='Empty model configuration'):
        ModelReader(_cfg(tmp_path, ''), [2], (4,)).build_model()
    with pytest.raises(AssertionError, match='Output units given are less than dense layers required'):
        ModelReader(ANN, [2], (4,)).build_model()
    with pytest.raises(AssertionError, match='Unsupported activation'):
        ModelReader(_cfg(tmp_path, '[dense-0]\nunits=3\nactivation=swishy\noutput=1\n'), [], (4,)).build_model()
    with pytest.raises(AssertionError, match='smaller than the kernel'):
        ModelReader(CNN, [2, 1], (84, 4, 4)).build_model()
    reader = ModelReader(ANN, [2, 1], (4,))
    reader.build_model()
    assert reader.output_count == 0                                                 # reusable, like the reference's


def test_tensor_core_dense_selection():
    """Dense layers move to the tcgen05 GEMM only where the kernel's operand constraints hold."""
    from xagents_b200.agents.tc_dense import TcLinear
    net = ModelReader(CNN, [6, 1], (84, 84, 4), conv_dims=2, tensor_core_dense=True).build_model()
    assert [isinstance(layer.linear, TcLinear) for layer in net.layers[4:]] == [True, True, True]
    ann = ModelReader(ANN, [2, 1], (4,), tensor_core_dense=True).build_model()
    assert [isinstance(layer.linear, TcLinear) for layer in ann.layers] == [False, False, True, True]  # K=4 / tanh stay


# ---------------------------------------------------------------------- environments
def test_cartpole_follows_the_published_dynamics():
    env = envs.CartPole(seed=5)
    s = env.reset()
    assert s.shape == (4,) and s.dtype == np.float32 and np.abs(s).max() <= 0.05
    x, xd, th, thd = env.state
    nxt, reward, done, info = env.step(1)
    total, pml = 1.1, 0.05
    temp = (10.0 + pml * thd * thd * math.sin(th)) / total
    th_acc = (9.8 * math.sin(th) - math.cos(th) * temp) / (0.5 * (4.0 / 3.0 - 0.1 * math.cos(th) ** 2 / total))
    x_acc = temp - pml * th_acc * math.cos(th) / total
    want = np.array([x + 0.02 * xd, xd + 0.02 * x_acc, th + 0.02 * thd, thd + 0.02 * th_acc], np.float32)
    assert np.allclose(nxt, want, rtol=1e-6) and reward == 1.0 and not done
    steps = 1
    while not done:                                                 # pushing one way only: the pole falls quickly
        nxt, reward, done, info = env.step(1)
        steps += 1
    assert steps < 30 and abs(nxt[2]) > env.THETA_LIMIT
    a, b = envs.CartPole(seed=9), envs.CartPole(seed=9)
    assert np.array_equal(a.reset(), b.reset())


def test_cartpole_time_limit():
    env = envs.CartPole(seed=0)
    env.reset()
    for t in range(500):                                            # a bang-bang controller on the pole angle keeps it up
        s, r, done, info = env.step(1 if env.state[2] + 0.5 * env.state[3] > 0 else 0)
        if done:
            break
    assert t == 499 and done and info['TimeLimit.truncated']


def test_synthetic_atari_and_create_envs():
    made = envs.create_envs('SyntheticAtari-v0', 3, preprocess=True)
    assert len(made) == 3 and made[0].observation_space.shape == (84, 84, 4) and made[0].action_space.n == 6
    frame = made[0].reset()
    assert frame.shape == (84, 84, 4) and frame.dtype == np.uint8
    made[0].seed(1)
    rewards, dones = zip(*[(made[0].step(0)[1], made[0].step(0)[2]) for _ in range(2000)])
    assert set(rewards) <= {-1.0, 0.0, 1.0} and 0 < sum(r != 0 for r in rewards) < 200 and 0 < sum(dones) < 100
    with pytest.raises(AssertionError, match='Cannot use AtariWrapper or --preprocess for non-atari environment CartPole-v1'):
        envs.create_envs('CartPole-v1', 1, preprocess=True)
    with pytest.raises(ImportError, match='neither gym nor gymnasium'):
        envs.create_envs('NoSuchEnv-v0', 1, preprocess=False)


# ---------------------------------------------------------------------- command line
def test_flag_tables_carry_the_reference_defaults():
    """Defaults of xagents/utils/cli.py and xagents/{a2c,ppo}/cli.py."""
    ex = cli.Executor()
    ex.command, ex.agent_id = 'train', 'ppo'
    agent, general, command = ex.parse_known_args(['train', 'ppo', '--env', 'CartPole-v1', '--max-steps', '10'])
    assert vars(command) == {'target_reward': None, 'max_steps': 10, 'monitor_session': None}
    want_agent = dict(reward_buffer_size=100, gamma=0.99, display_precision=2, seed=None, log_frequency=None, checkpoints=None,
                      history_checkpoint=None, plateau_reduce_factor=0.9, plateau_reduce_patience=10, early_stop_patience=3,
                      divergence_monitoring_steps=None, quiet=None, model=None, entropy_coef=0.01, value_loss_coef=0.5,
                      grad_norm=0.5, n_steps=128, lam=0.95, ppo_epochs=4, mini_batches=4, advantage_epsilon=1e-8, clip_norm=0.1)
    assert vars(agent) == want_agent
    g = vars(general)
    assert (g['env'], g['n_envs'], g['lr'], g['opt_epsilon'], g['beta1'], g['beta2'], g['weights']) == \
        ('CartPole-v1', 1, 7e-4, 1e-7, 0.9, 0.999, None)
    ex.agent_id = 'a2c'
    agent, _, _ = ex.parse_known_args(['train', 'a2c', '--env', 'x', '--target-reward', '3', '--n-steps', '7', '--quiet'])
    assert agent.n_steps == 7 and agent.quiet is True and not hasattr(agent, 'lam')
    assert cli.a2c_args['n-steps']['default'] == 5


def test_command_line_errors_and_help(capsys):
    with pytest.raises(AssertionError, match='Invalid command `fly`'):
        cli.execute(['fly'])
    with pytest.raises(AssertionError, match='Invalid agent `dqn`'):       # off-policy agents are outside the hot path
        cli.execute(['train', 'dqn'])
    with pytest.raises(AssertionError, match='train requires --target-reward or --max-steps'):
        cli.execute(['train', 'ppo', '--env', 'CartPole-v1'])
    cli.Executor().execute([])
    assert 'Available commands' in capsys.readouterr().out
    cli.execute(['train', 'ppo'])
    out = capsys.readouterr().out
    for flag in ('--lam', '--mini-batches', '--clip-norm', '--env', '--max-steps', '--grad-norm', '--conv-dims'):
        assert flag in out
    ex = cli.Executor()
    ex.command, ex.agent_id = 'train', 'a2c'
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter('always')
        ex.parse_known_args(['train', 'a2c', '--env', 'x', '--max-steps', '5', '--no-such-flag'])
    assert any('--no-such-flag' in str(w.message) for w in caught)


def test_trpo_flags_and_default_models():
    """xagents/trpo/cli.py: PPO's flags plus TRPO's, `model` removed, different defaults for entropy-coef / lam / n-steps."""
    ex = cli.Executor()
    ex.command, ex.agent_id = 'train', 'trpo'
    agent, _, _ = ex.parse_known_args(['train', 'trpo', '--env', 'CartPole-v1', '--max-steps', '10'])
    a = vars(agent)
    assert 'model' not in a and a['actor_model'] is None and a['critic_model'] is None
    assert (a['max_kl'], a['cg_iterations'], a['cg_residual_tolerance'], a['cg_damping'], a['actor_iterations'],
            a['critic_iterations'], a['fvp_n_steps'], a['entropy_coef'], a['lam'], a['n_steps']) == \
        (1e-3, 10, 1e-10, 1e-3, 10, 3, 5, 0, 1.0, 512)
    assert (a['ppo_epochs'], a['mini_batches'], a['clip_norm'], a['grad_norm']) == (4, 4, 0.1, 0.5)
    assert cli.ppo_args['lam']['default'] == 0.95 and 'model' in cli.ppo_args          # the copies are independent
    reg = cli.agents['trpo']
    assert 'model' not in reg
    assert [os.path.basename(p) for p in reg['actor_model']['ann'] + reg['actor_model']['cnn']] == ['ann-actor.cfg', 'cnn-actor.cfg']
    assert [os.path.basename(p) for p in reg['critic_model']['ann'] + reg['critic_model']['cnn']] == ['ann-critic.cfg', 'cnn-critic.cfg']
    actor = ModelReader(reg['actor_model']['cnn'][0], [6], (84, 84, 4)).build_model()
    critic = ModelReader(reg['critic_model']['ann'][0], [1], (4,)).build_model()
    assert actor(torch.rand(2, 84, 84, 4)).shape == (2, 6) and critic(torch.rand(3, 4)).shape == (3, 1)
    assert sum(p.numel() for p in actor.parameters()) == 28224 * 128 + 128 + 128 * 6 + 6


def test_registry_shape():
    """xagents.agents[id] keeps its keys (xagents/__init__.py:18-27; register_models, common.py:312-343)."""
    from xagents_b200.agents import A2C, PPO
    assert cli.agents['ppo']['agent'] is PPO and cli.agents['a2c']['agent'] is A2C
    for agent_id in ('a2c', 'ppo'):
        assert [os.path.basename(p) for p in cli.agents[agent_id]['model']['cnn']] == ['cnn-actor-critic.cfg']
        assert [os.path.basename(p) for p in cli.agents[agent_id]['model']['ann']] == ['ann-actor-critic.cfg']
        assert 'n-steps' in cli.agents[agent_id]['module'].cli_args


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU behaviour')
def test_create_agent_fails_loudly_without_a_gpu():
    """No CPU fallback: building the agent needs the device."""
    with pytest.raises((RuntimeError, AssertionError)):
        cli.execute(['train', 'ppo', '--env', 'CartPole-v1', '--n-envs', '2', '--max-steps', '10', '--quiet'])


def test_replay_buffer_handles_and_acer_flags():
    """xagents/utils/buffers.py:8-60 assertions; create_buffers' as_total split (common.py:497-565); xagents/acer/cli.py."""
    from xagents_b200.buffers import BaseBuffer, ReplayBuffer1, create_buffers
    for kwargs, message in ((dict(size=0), 'Buffer size should be > 0'), (dict(size=4, initial_size=0), 'Buffer initial size should be > 0'),
                            (dict(size=4, batch_size=0), 'Buffer batch size should be > 0'),
                            (dict(size=4, batch_size=5), 'should be <= size'), (dict(size=4, initial_size=5, batch_size=1), 'exceeds max size')):
        with pytest.raises(AssertionError, match=message):
            BaseBuffer(**kwargs)
    with pytest.raises(NotImplementedError):
        BaseBuffer(4, batch_size=1).get_sample()
    made = create_buffers('acer', 10000, 32, 16, None)
    assert len(made) == 16 and all(isinstance(b, ReplayBuffer1) for b in made)
    assert (made[0].size, made[0].initial_size, made[0].batch_size, made[0].current_size) == (625, 625, 1, 0)
    assert create_buffers('acer', 100, 32, 4, 20, as_total=False)[0].initial_size == 20
    ex = cli.Executor()
    ex.command, ex.agent_id = 'train', 'acer'
    agent, general, _ = ex.parse_known_args(['train', 'acer', '--env', 'x', '--max-steps', '10'])
    a, g = vars(agent), vars(general)
    assert (a['ema_alpha'], a['replay_ratio'], a['epsilon'], a['importance_c'], a['delta'], a['trust_region'], a['n_steps'],
            a['grad_norm'], a['entropy_coef']) == (0.99, 4, 1e-6, 10.0, 1, None, 20, 10, 0.01)
    assert (g['buffer_max_size'], g['buffer_initial_size'], g['buffer_batch_size']) == (10000, None, 32)
    ex.agent_id = 'ppo'
    _, general, _ = ex.parse_known_args(['train', 'ppo', '--env', 'x', '--max-steps', '10'])
    assert 'buffer_max_size' not in vars(general)                   # buffer flags only for agents that own a buffer
    net = ModelReader(cli.agents['acer']['model']['cnn'][0], [6, 6], (84, 84, 4), conv_dims=2).build_model()
    probs, q = net(torch.rand(2, 84, 84, 4))
    assert net.output_is_softmax and probs.shape == q.shape == (2, 6) and torch.allclose(probs.sum(-1), torch.ones(2), atol=1e-6)


def test_batched_device_environment_keeps_the_step_envs_contract():
    """envs.BatchedSyntheticAtari behind BaseAgent.step_envs (xagents/base.py:388-426): same outputs and bookkeeping as the
    per-environment loop -- terminal frame returned, post-reset frame kept, episode sums cut at dones -- as tensors."""
    from xagents_b200.agents import BaseAgent
    made = envs.create_envs('SyntheticAtariDevice-v0', 6, preprocess=True, device='cpu')
    assert made.batched and len(made) == 6 and made[3].observation_space.shape == (84, 84, 4)
    with pytest.raises(TypeError, match='not a list of environments'):
        list(made)
    made.p_done, made.p_reward = 0.3, 0.6
    agent = BaseAgent(made, None, n_steps=4, quiet=True, seed=4, device='cpu')
    assert agent.batched and agent.n_envs == 6 and agent.img_inputs and agent.n_actions == 6
    assert isinstance(agent.get_states(), torch.Tensor) and agent.get_states().dtype == torch.uint8
    sums, finished, n_done = np.zeros(6), [], 0
    for step in range(25):
        before = agent.get_states().clone()
        previous, actions, rewards, dones, new_states = agent.step_envs(torch.zeros(6), True)
        assert torch.equal(previous, before) and torch.equal(agent.get_dones(), dones)
        assert set(rewards.tolist()) <= {-1.0, 0.0, 1.0} and set(dones.tolist()) <= {0.0, 1.0}
        moved_on = (agent.get_states() != new_states).flatten(1).any(1)
        assert torch.equal(moved_on, dones.bool())                 # finished envs already hold their post-reset frame
        sums += rewards.numpy()
        for e in np.nonzero(dones.numpy())[0]:
            finished.append(sums[e])
            sums[e] = 0
            n_done += 1
    assert agent.steps == 25 * 6 and agent.games == 0              # bookkeeping is deferred ...
    from time import perf_counter
    agent.training_start_time = agent.last_reset_time = perf_counter()      # what fit() sets before its loop
    agent.check_episodes()                                         # ... to one read-back per train step
    assert agent.games == n_done > 10 and list(agent.total_rewards) == [float(x) for x in finished][-100:]
    assert np.allclose(agent._episode_sums.numpy(), sums)
    agent.update_metrics()
    assert agent.mean_reward == pytest.approx(np.around(np.mean(finished[-100:]), 2))      # rounded like base.py:289-291
    assert agent.step_envs(torch.zeros(6)) == []


def test_batched_cartpole_integrates_like_the_per_environment_cartpole():
    """envs.BatchedCartPole against envs.CartPole from identical states under identical actions: same trajectories,
    same termination steps, time limit included; finished episodes restart inside the initial-state box."""
    made = envs.create_envs('CartPoleDevice-v1', 5, preprocess=False, device='cpu')
    made.seed(11)
    first = made.reset_all()
    assert first.shape == (5, 4) and first.dtype == torch.float32 and first.abs().max() <= 0.05
    singles = [envs.CartPole() for _ in range(5)]
    for i, env in enumerate(singles):
        env.reset()
        env.state = made.state[i].numpy().copy()
    rng = np.random.default_rng(0)
    alive, finished = [True] * 5, 0
    for step in range(120):
        actions = rng.integers(0, 2, 5)
        new_states, rewards, dones = made.step_all(torch.as_tensor(actions, dtype=torch.float32))
        assert torch.equal(rewards, torch.ones(5))
        for i, env in enumerate(singles):
            if not alive[i]:
                continue
            want, reward, done, _ = env.step(int(actions[i]))
            assert np.allclose(new_states[i].numpy(), want, rtol=0, atol=1e-6) and bool(dones[i]) == done
            if done:                                                # the batched env restarted it; stop following this one
                alive[i], finished = False, finished + 1
                assert made.states[i].abs().max() <= 0.05 and int(made.t[i]) == 0
                assert not torch.equal(made.states[i], new_states[i])
            else:
                assert torch.equal(made.states[i], new_states[i])
    assert finished == 5                                            # random actions drop every pole well within 120 steps
    made.reset_all()                                                # time limit: a bang-bang controller holds for 500 steps
    for t in range(500):
        s = made.state
        _, _, dones = made.step_all((s[:, 2] + 0.5 * s[:, 3] > 0).float())
        assert bool(dones.any()) == (t == 499)
    assert bool(dones.all())
    with pytest.raises(AssertionError, match='Cannot use AtariWrapper or --preprocess for non-atari environment CartPoleDevice-v1'):
        envs.create_envs('CartPoleDevice-v1', 2, preprocess=True, device='cpu')


def test_step_envs_columns_for_a_list_of_environments():
    """BaseAgent.step_envs(get_observation=True) -> [state, action, reward, done, new_state] (xagents/base.py:388-426): the
    terminal frame is returned while the agent already holds the reset frame; uint8 frames stay uint8, everything else is
    float32 like the reference's columns."""
    from xagents_b200.agents import BaseAgent
    made = envs.create_envs('SyntheticAtari-v0', 5, preprocess=True)
    for i, env in enumerate(made):
        env.seed(i)
        env.p_done = 0.4
    agent = BaseAgent(made, None, n_steps=3, quiet=True, device='cpu')
    finished = 0
    for _ in range(12):
        before = agent.get_states()
        state, action, reward, done, new_state = agent.step_envs(np.arange(5) % 6, True)
        assert [c.dtype for c in (state, action, reward, done, new_state)] == [np.uint8, np.float32, np.float32, np.float32, np.uint8]
        assert np.array_equal(state, before) and np.array_equal(action, np.arange(5, dtype=np.float32))
        held = agent.get_states()
        for e in range(5):
            assert bool(done[e]) == agent.dones[e]
            if done[e]:
                finished += 1
            else:
                assert np.array_equal(held[e], new_state[e])
    assert agent.games == finished > 5 and agent.steps == 60 and len(agent.total_rewards) == finished
    vector = BaseAgent(envs.create_envs('CartPole-v1', 3, preprocess=False), None, quiet=True, device='cpu')
    assert [c.dtype for c in vector.step_envs(np.zeros(3, np.int64), True)] == [np.float32] * 5


def test_plateau_learning_rate_reduction_and_early_stop_follow_the_reference():
    """base.py:213-230, 270-291, 333-335: under `divergence_monitoring_steps` a mean reward that does not beat the best one
    counts towards a plateau; `plateau_reduce_patience` plateaus scale the learning rate by `plateau_reduce_factor` and count
    towards early stopping; a new best reward resets both counters."""
    import types

    from xagents_b200.agents import BaseAgent
    made = envs.create_envs('SyntheticAtariDevice-v0', 2, preprocess=True, device='cpu')
    model = types.SimpleNamespace(lr=1e-3)
    agent = BaseAgent(made, model, n_steps=4, quiet=True, device='cpu', divergence_monitoring_steps=10, plateau_reduce_factor=0.5,
                      plateau_reduce_patience=2, early_stop_patience=2)
    agent.max_steps = 10 ** 9
    agent.steps = 5
    agent.total_rewards.extend([1.0, 3.0])
    agent.update_metrics()                                         # below divergence_monitoring_steps: nothing counts
    assert (agent.plateau_count, agent.early_stop_count, agent.mean_reward, model.lr) == (0, 0, 2.0, 1e-3)
    agent.steps = 20
    agent.update_metrics()                                         # mean 2.0 becomes the best reward: counters reset
    assert agent.best_reward == 2.0 and agent.plateau_count == 1   # ... and mean <= best already counts (the reference's order)
    agent.update_metrics()
    assert agent.plateau_count == 0 and agent.early_stop_count == 1 and model.lr == pytest.approx(5e-4)
    assert not agent.training_done()
    agent.update_metrics()
    agent.update_metrics()
    assert agent.early_stop_count == 2 and model.lr == pytest.approx(2.5e-4) and agent.training_done()
    agent.total_rewards.extend([50.0] * 100)                       # a better mean arrives: both counters start over
    agent.update_metrics()                                         # (this call still sees the old mean ...)
    agent.update_metrics()                                         # (... the next one the new best)
    assert agent.best_reward == 50.0 and agent.early_stop_count == 0 and not agent.training_done()
