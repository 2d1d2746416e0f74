"""`xagents train <a2c|acer|ppo|trpo> ...` on the device path: the reference's command line for the on-policy agents this
package mirrors (xagents/cli.py:13-241, xagents/utils/cli.py, xagents/{a2c,ppo}/cli.py) and its factory
(`create_model(s)` / `create_agent`, xagents/utils/common.py:430-494, 568-624).

    python -m xagents_b200 train ppo --env CartPole-v1 --n-envs 16 --max-steps 200000
    python -m xagents_b200 train a2c --env SyntheticAtari-v0 --n-envs 16 --max-steps 4000 --conv-dims 2

Flags, defaults, the agent/non-agent/command split and the assertion messages are the reference's.  `play`, `tune`
and the off-policy agents (dqn, ddpg, td3) are outside the hot path and are refused as invalid commands /
agents;
checkpoint, history and wandb flags are accepted by the parser and refused by the agent (agents/base.py).
Four flags are new: `--conv-dims` (1 = the `.cfg` as the reference's reader builds it, Conv1D; 2 = the documented
Conv2D network), `--tensor-core-dense` (Dense layers on the tcgen05 GEMM), `--tensor-core-network` (that Conv2D network, forward
and backward, on the tcgen05 kernels: `agents.NatureCnnTc` with the file's initialisers and seed) and `--device`.
"""
import argparse
import sys
import types
import warnings
from pathlib import Path

from . import __version__, buffers as _buffers, envs as _envs
from .agents import A2C, ACER, PPO, TRPO
from .agents.cfg import ModelReader

_MODELS = Path(__file__).parent / 'agents' / 'models'


def _flag(help, type=None, default=None, action=None, required=None, nargs=None):
    spec = {'help': help}
    for key, value in (('type', type), ('default', default), ('action', action), ('required', required), ('nargs', nargs)):
        if value is not None:
            spec[key] = value
    return spec


non_agent_args = {
    'env': _flag('environment id (built-in: CartPole-v1, SyntheticAtari-v0, SyntheticAtariDevice-v0; else gym / gymnasium)',
                 required=True),
    'n-envs': _flag('Number of environments to create', int, 1),
    'preprocess': _flag('Treat states as atari frames and preprocess them', action='store_true'),
    'lr': _flag('Adam learning rate', float, 7e-4),
    'opt-epsilon': _flag('Adam epsilon', float, 1e-7),
    'beta1': _flag('Adam beta1', float, 0.9),
    'beta2': _flag('Adam beta2', float, 0.999),
    'weights': _flag('Path(s) to model weights (torch state_dict) loaded into the agent output_models', nargs='+'),
    'max-frame': _flag('Max & skip during preprocessing', action='store_true'),
    'conv-dims': _flag('1: convolutional sections as the reference reader builds them (Conv1D); 2: Conv2D', int, 1),
    'tensor-core-dense': _flag('Dense layers on the tcgen05 GEMM', action='store_true'),
    'tensor-core-network': _flag('the documented Conv2D actor-critic CNN, forward and backward, on the tcgen05 kernels (a2c / ppo, 84x84x4 frames)',
                                 action='store_true'),
    'device': _flag('CUDA device of this process', default='cuda:0'),
}

off_policy_args = {
    'buffer-max-size': _flag('Maximum replay buffer size', int, 10000),
    'buffer-initial-size': _flag('Replay buffer initial size', int),
    'buffer-batch-size': _flag('Replay buffer batch size', int, 32),
}

agent_args = {
    'reward-buffer-size': _flag('Size of the total reward buffer used for the displayed mean reward', int, 100),
    'gamma': _flag('Discount factor', float, 0.99),
    'display-precision': _flag('Number of decimals to be displayed', int, 2),
    'seed': _flag('Random seed', int),
    'log-frequency': _flag('Log progress every n games', int),
    'checkpoints': _flag('Path(s) to which checkpoint(s) would be saved (refused: outside the hot path)', nargs='+'),
    'history-checkpoint': _flag('Path to .parquet training history (refused: outside the hot path)'),
    'plateau-reduce-factor': _flag('Factor multiplied by the learning rate on a plateau', float, 0.9),
    'plateau-reduce-patience': _flag('Minimum non-improvements to reduce lr', int, 10),
    'early-stop-patience': _flag('Minimum plateau reduces to stop training', int, 3),
    'divergence-monitoring-steps': _flag('Steps after which plateau and early stopping are active', int),
    'quiet': _flag('No messages by the agent', action='store_true'),
}

train_args = {
    'target-reward': _flag('Target reward: training stops when reached', int),
    'max-steps': _flag('Maximum number of environment steps', int),
    'monitor-session': _flag('Wandb session name (refused: outside the hot path)'),
}

a2c_args = {
    'model': _flag('Path to model .cfg file'),
    'entropy-coef': _flag('Entropy coefficient of the loss', float, 0.01),
    'value-loss-coef': _flag('Value loss coefficient of the loss', float, 0.5),
    'grad-norm': _flag('Global-norm gradient clipping value', float, 0.5),
    'n-steps': _flag('Transition steps', int, 5),
}

ppo_args = dict(a2c_args)
ppo_args.update({
    'lam': _flag('GAE-Lambda for advantage estimation', float, 0.95),
    'ppo-epochs': _flag('Gradient updates per training step', int, 4),
    'mini-batches': _flag('Number of mini-batches per update', int, 4),
    'advantage-epsilon': _flag('Value added to the advantage standard deviation', float, 1e-8),
    'clip-norm': _flag('Clipping value of the probability ratio and of the value update', float, 0.1),
    'n-steps': _flag('Transition steps', int, 128),
})

acer_args = dict(a2c_args)
acer_args.update({
    'ema-alpha': _flag('Decay of the averaged policy network', float, 0.99),
    'replay-ratio': _flag('Mean of the Poisson number of replay updates per train step', int, 4),
    'epsilon': _flag('Epsilon used in several calculations during the gradient update', float, 1e-6),
    'importance-c': _flag('Importance weight truncation parameter', float, 10.0),
    'delta': _flag('Delta parameter of the trust region update', float, 1),
    'trust-region': _flag('If specified, trust region updates will be used', action='store_true'),
    'n-steps': _flag('Transition steps', int, 20),
    'grad-norm': _flag('Global-norm gradient clipping value', float, 10),
})

trpo_args = dict(ppo_args)
trpo_args.update({
    'actor-model': _flag('Path to actor model .cfg file'),
    'critic-model': _flag('Path to critic model .cfg file'),
    'max-kl': _flag('Maximum KL divergence of an actor update', float, 1e-3),
    'cg-iterations': _flag('Conjugate-gradient iterations per train step', int, 10),
    'cg-residual-tolerance': _flag('Conjugate-gradient residual tolerance', float, 1e-10),
    'cg-damping': _flag('Damping added to the Fisher-vector product', float, 1e-3),
    'actor-iterations': _flag('Line-search iterations per train step', int, 10),
    'critic-iterations': _flag('Critic optimisation passes per train step', int, 3),
    'fvp-n-steps': _flag('Use every n-th state for Fisher-vector products', int, 5),
    'entropy-coef': _flag('Entropy coefficient of the loss', float, 0),
    'lam': _flag('GAE-Lambda for advantage estimation', float, 1.0),
    'n-steps': _flag('Transition steps', int, 512),
})
del trpo_args['model']


def _default_models(role, folder=_MODELS):
    """Default `.cfg` files by network type, split like register_models (common.py:312-343): files naming both `actor`
    and `critic` are single-model defaults, the others belong to the role they name."""
    groups = {'cnn': [], 'ann': []}
    for cfg in sorted(folder.iterdir()):
        has = {'actor': 'actor' in cfg.name, 'critic': 'critic' in cfg.name}
        mine = (has['actor'] and has['critic']) if role == 'model' else (has[role.split('_')[0]] and sum(has.values()) == 1)
        for kind in groups:
            if cfg.suffix == '.cfg' and kind in cfg.name and mine:
                groups[kind].append(cfg.as_posix())
    return groups


# xagents.agents / xagents.commands (xagents/__init__.py:18-40), restricted to what this package mirrors
agents = {
    'a2c': {'module': types.SimpleNamespace(cli_args=a2c_args, __file__=str(_MODELS.parent / 'a2c.py')), 'agent': A2C,
            'model': _default_models('model')},
    'acer': {'module': types.SimpleNamespace(cli_args=acer_args, __file__=str(_MODELS.parent / 'acer.py')), 'agent': ACER,
             'model': _default_models('model', _MODELS / 'acer')},
    'ppo': {'module': types.SimpleNamespace(cli_args=ppo_args, __file__=str(_MODELS.parent / 'ppo.py')), 'agent': PPO,
            'model': _default_models('model')},
    'trpo': {'module': types.SimpleNamespace(cli_args=trpo_args, __file__=str(_MODELS.parent / 'trpo.py')), 'agent': TRPO,
             'actor_model': _default_models('actor_model'), 'critic_model': _default_models('critic_model')},
}
commands = {'train': (train_args, 'fit', 'Train given an agent and environment')}


# ---------------------------------------------------------------------- factory
def create_model(env, agent_id, model_type, optimizer_kwargs=None, seed=None, model_cfg=None, conv_dims=1,
                 tensor_core_dense=False, device='cuda:0', tensor_core_network=False):
    """`.cfg` -> network -> device adapter.  Head widths as in common.py:447-471: n_actions (Discrete) or the action
    dimension (Box), then 1 for an actor-critic file."""
    space = env.action_space
    units = [int(space.n) if hasattr(space, 'n') else int(space.shape[0])]
    network_type = 'cnn' if len(env.observation_space.shape) == 3 else 'ann'
    try:
        model_cfg = model_cfg or agents[agent_id][model_type][network_type][0]
    except (IndexError, KeyError):
        model_cfg = None
    assert model_cfg, (f'You should specify `model_cfg`. No default {network_type.upper()} model found in\n{_MODELS}')
    name = Path(model_cfg).name
    if agent_id == 'acer':                                         # policy head + one critic output per action
        units.append(units[-1])
    elif 'actor' in name and 'critic' in name:
        units.append(1)
    elif 'critic' in name:
        units[0] = 1
    role = {'model': 'actor_critic', 'actor_model': 'actor', 'critic_model': 'critic'}[model_type]
    if tensor_core_network:
        assert role == 'actor_critic' and agent_id in ('a2c', 'ppo'), '--tensor-core-network is the actor-critic CNN of a2c / ppo'
        reader = ModelReader(model_cfg, units, env.observation_space.shape, optimizer_kwargs or {}, seed, conv_dims=2)
        return reader.build_adapter(device, role=role, tensor_core_network=True)
    reader = ModelReader(model_cfg, units, env.observation_space.shape, optimizer_kwargs or {}, seed, conv_dims=conv_dims,
                         tensor_core_dense=tensor_core_dense and role != 'actor')
    return reader.build_adapter(device, role=role)


def create_models(options, env, agent_id, **kwargs):
    models = {}
    for model_type in ('model', 'actor_model', 'critic_model'):
        if model_type in options:
            model_cfg = options[model_type]
            if not isinstance(model_cfg, (str, Path)):
                model_cfg = None
            models[model_type] = create_model(env, agent_id, model_type, model_cfg=model_cfg, **kwargs)
    return models


def _rank_device(local_rank):
    return f'cuda:{local_rank}'


def create_agent(agent_id, agent_kwargs, non_agent_kwargs, trial=None):
    """Environments + model + agent from parsed flags (common.py:568-624).

    Under `torchrun` (WORLD_SIZE > 1) the job shards its environments: `--n-envs` is the job-wide count, rank g builds
    and steps only its slice on `cuda:LOCAL_RANK`, the networks start from rank 0's weights and every update averages
    gradients over the ranks (`dist.ShardComm`, collectives C1 / C2 of SURVEY.md 8e).  The reference is single-process;
    there is nothing to mirror, only to keep: the flags mean what they meant."""
    import torch
    from . import dist as xdist
    assert agent_id in agents, f'Invalid agent `{agent_id}`'
    agent_kwargs = dict(agent_kwargs)
    if trial is not None:
        agent_kwargs['trial'] = trial
    device = non_agent_kwargs.get('device') or 'cuda:0'
    rank, local_rank, world = xdist.init_from_env()
    comm, n_envs = None, non_agent_kwargs['n_envs']
    if world > 1:
        assert agent_id in ('a2c', 'ppo'), (f'`{agent_id}` cannot shard its environments: its updates are not a fixed '
                                            f'sequence of gradient steps (line search / random replay counts)')
        assert n_envs % world == 0, f'--n-envs {n_envs} must be a multiple of the {world} ranks'
        device = _rank_device(local_rank)
        comm = xdist.ShardComm(device=device)
        n_envs //= world
        if rank > 0:
            agent_kwargs['quiet'] = True
        if agent_kwargs.get('seed') is not None:                   # different episodes on every rank, same model seed
            non_agent_kwargs = dict(non_agent_kwargs, env_seed=agent_kwargs['seed'] + 7919 * rank)
    envs = _envs.create_envs(non_agent_kwargs['env'], n_envs, non_agent_kwargs['preprocess'],
                             max_frame=non_agent_kwargs.get('max_frame'), device=device)
    agent_kwargs['envs'] = envs
    optimizer_kwargs = {'learning_rate': non_agent_kwargs['lr'], 'beta_1': non_agent_kwargs['beta1'],
                        'beta_2': non_agent_kwargs['beta2'], 'epsilon': non_agent_kwargs['opt_epsilon']}
    models = create_models(agent_kwargs, envs[0], agent_id, optimizer_kwargs=optimizer_kwargs,
                           seed=agent_kwargs.get('seed'), conv_dims=non_agent_kwargs.get('conv_dims', 1),
                           tensor_core_dense=bool(non_agent_kwargs.get('tensor_core_dense')), device=device,
                           tensor_core_network=bool(non_agent_kwargs.get('tensor_core_network')))
    if comm is not None:
        for model in models.values():
            comm.broadcast_(model.flat_param)
            for mod in model._refreshable:
                mod.refresh()
            model.comm = comm
    agent_kwargs.update(models)
    if agent_id == 'acer':
        agent_kwargs['buffers'] = _buffers.create_buffers(agent_id, non_agent_kwargs['buffer_max_size'],
                                                          non_agent_kwargs['buffer_batch_size'], non_agent_kwargs['n_envs'],
                                                          non_agent_kwargs['buffer_initial_size'])
    agent = agents[agent_id]['agent'](device=device, **agent_kwargs)
    if comm is not None:
        agent._rng_offset = rank << 40                             # disjoint Philox counter ranges for the action sampler
        if 'env_seed' in non_agent_kwargs:
            if getattr(envs, 'batched', False):
                envs.seed(non_agent_kwargs['env_seed'])
            else:
                for i, env in enumerate(envs):
                    env.seed(non_agent_kwargs['env_seed'] + i)
            agent.reset_envs()
    if non_agent_kwargs.get('weights'):
        n_weights, n_models = len(non_agent_kwargs['weights']), len(agent.output_models)
        assert n_weights == n_models, f'Expected {n_models} weights to load, got {n_weights}'
        for weight, model in zip(non_agent_kwargs['weights'], agent.output_models):
            model.module.load_state_dict(torch.load(weight, map_location=device))
            for mod in getattr(model, '_refreshable', []):
                mod.refresh()
    return agent


# ---------------------------------------------------------------------- command line
class Executor:
    def __init__(self):
        self.agent_id = None
        self.command = None
        self.agent = None

    @staticmethod
    def display_section(title, cli_args):
        print(f'\n{title}\n')
        width = max(len(flag) for flag in cli_args) + 2
        for flag in sorted(cli_args):
            spec = cli_args[flag]
            default = spec.get('default', '-')
            print(f'  --{flag:<{width}} {spec["help"]}  [default: {default}]')

    def display_commands(self, sections=None):
        print(f'xagents_b200 {__version__}')
        print('\nUsage:')
        print('\txagents <command> <agent> [options] [args]')
        print('\nAvailable commands:')
        for command, items in commands.items():
            print(f'\t{command:<10} {items[2]}')
        print()
        print('Use xagents <command> to see more info about a command')
        print('Use xagents <command> <agent> to see more info about command + agent')
        for title, cli_args in (sections or {}).items():
            self.display_section(title, cli_args)

    @staticmethod
    def add_args(cli_args, parser):
        for arg, options in cli_args.items():
            if options.get('action'):
                parser.add_argument(f'--{arg}', help=options.get('help'), default=options.get('default'),
                                    action=options['action'])
            else:
                parser.add_argument(f'--{arg}', help=options.get('help'), default=options.get('default'),
                                    type=options.get('type'), required=options.get('required'), nargs=options.get('nargs'))

    def maybe_create_agent(self, argv):
        to_display = {}
        if len(argv) == 0:
            self.display_commands()
            return
        command = argv[0]
        to_display.update(non_agent_args)
        to_display.update(agent_args)
        assert command in commands, f'Invalid command `{command}`'
        to_display.update(commands[command][0])
        if len(argv) == 1:
            self.display_commands({command: to_display})
            return
        agent_id = argv[1]
        assert agent_id in agents, f'Invalid agent `{agent_id}`'
        to_display.update(agents[agent_id]['module'].cli_args)
        if len(argv) == 2:
            if agent_id == 'acer':
                to_display.update(off_policy_args)
            self.display_commands({f'{command} {agent_id}': to_display})
            return
        self.command, self.agent_id = command, agent_id

    def parse_known_args(self, argv):
        general_parser, agent_parser, command_parser = (argparse.ArgumentParser(add_help=False) for _ in range(3))
        self.add_args(agent_args, agent_parser)
        self.add_args(agents[self.agent_id]['module'].cli_args, agent_parser)
        self.add_args(commands[self.command][0], command_parser)
        if self.agent_id == 'acer':
            self.add_args(off_policy_args, general_parser)
        self.add_args(non_agent_args, general_parser)
        non_agent_known, extra1 = general_parser.parse_known_args(argv)
        agent_known, extra2 = agent_parser.parse_known_args(argv)
        command_known, extra3 = command_parser.parse_known_args(argv)
        unknown_flags = [flag for flag in set(extra1) & set(extra2) & set(extra3)
                         if flag not in (self.command, self.agent_id) and '--' in flag]
        if unknown_flags:
            warnings.warn(f'Got unknown flags {unknown_flags}')
        if self.command == 'train':
            assert command_known.target_reward or command_known.max_steps, 'train requires --target-reward or --max-steps'
        return agent_known, non_agent_known, command_known

    def execute(self, argv):
        self.maybe_create_agent(argv)
        if not self.agent_id:
            return
        agent_known, non_agent_known, command_known = self.parse_known_args(argv)
        self.agent = create_agent(self.agent_id, vars(agent_known), vars(non_agent_known))
        getattr(self.agent, commands[self.command][1])(**vars(command_known))


def execute(argv=None):
    argv = argv or sys.argv[1:]
    Executor().execute(argv)


if __name__ == '__main__':
    execute()
