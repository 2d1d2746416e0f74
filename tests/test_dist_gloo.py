"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo process group (no kernels here --
the arithmetic on each "rank" is the oracle's; what is under test is sharding, the equivalent global
permutation, and the two collectives as dist.ShardComm issues them)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from xagents_b200 import dist as xdist
from xagents_b200 import synthetic

T, E, K, MB = 12, 10, 2, 4


def test_shard_ranges_cover_envs_exactly_once():
    for n_envs in (1, 7, 16, 4096):
        for world in (1, 2, 3, 8):
            spans = [xdist.shard_range(n_envs, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n_envs
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    x = np.arange(6 * 10 * 3).reshape(6, 10, 3)
    assert np.array_equal(np.concatenate([xdist.shard_time_major(x, r, 2) for r in range(2)], axis=1), x)
    assert np.array_equal(np.concatenate([xdist.shard_time_major(np.arange(10), r, 2) for r in range(2)]), np.arange(10))


def test_combine_moments_is_the_whole_batch_statistics():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(1000) * 3 + 5
    parts = []
    for lo, hi in ((0, 100), (100, 100), (100, 640), (640, 1000)):
        seg = x[lo:hi]
        m = seg.mean() if len(seg) else 0.0
        parts.append((len(seg), m, ((seg - m) ** 2).sum() if len(seg) else 0.0, 0.0))
    n, mean, std = xdist.combine_moments(parts)
    assert n == 1000 and abs(mean - x.mean()) < 1e-12 and abs(std - x.std()) < 1e-12


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, lr, w = xdist.init_from_env(backend='gloo')
    assert (r, w) == (rank, world)
    comm = xdist.ShardComm(device=None)
    ro = synthetic.make_rollout(T, E, obs_shape=(4,), obs_dtype='float32', epochs=0, p_done=0.1)
    lo, hi = xdist.shard_range(E, rank, world)
    # local arithmetic (oracle stands in for the kernels on CPU): GAE needs no communication
    local = {k: xdist.shard_time_major(getattr(ro, k), rank, world) for k in ('rewards', 'values', 'dones', 'last_values')}
    returns = oracle.gae_returns(local['rewards'], local['dones'], local['values'], local['last_values'], 0.99, 0.95)
    n_local = (hi - lo) * T
    b_local = n_local // MB
    prng = np.random.default_rng(100 + rank)
    perms = [prng.permutation(n_local).astype(np.int32) for _ in range(K)]
    flat_ret, flat_val = oracle.concat_step_batches(returns, local['values'])
    # C2: local (count, mean, M2) of every minibatch -> all-gather -> combined statistics
    moments = torch.zeros((K * MB, 4), dtype=torch.float64)
    for k in range(K):
        for m in range(MB):
            idx = perms[k][m * b_local:(m + 1) * b_local]
            adv = (flat_ret[idx] - flat_val[idx]).astype(np.float64)
            moments[k * MB + m] = torch.tensor([len(adv), adv.mean(), ((adv - adv.mean()) ** 2).sum(), 0.0])
    gathered = torch.zeros((world, K * MB, 4), dtype=torch.float64)
    comm.all_gather_moments(gathered, moments)
    # C1: gradient all-reduce (sum), then 1/G
    grads = torch.full((1000,), float(rank + 1))
    comm.all_reduce_gradients_async(grads)
    comm.wait_gradients()
    slowest = comm.max_over_ranks(10.0 * (rank + 1))
    # job-wide stop conditions of a sharded fit() (agents/base.py: training_done) and the initial weight broadcast
    from collections import deque
    from xagents_b200.agents.base import BaseAgent
    agent = object.__new__(BaseAgent)
    agent.comm, agent.quiet, agent.batched = comm, True, False
    agent.early_stop_count, agent.early_stop_patience = 0, 3
    agent.total_rewards = deque([10.0, 20.0] if rank == 0 else [60.0], maxlen=100)
    agent.mean_reward = float(np.mean(agent.total_rewards))
    agent.steps = 100 * (rank + 1)
    decisions = []
    for target, max_steps in ((25.0, None), (31.0, None), (None, 300), (None, 301)):
        agent.target_reward, agent.max_steps = target, max_steps
        decisions.append(agent.training_done())
    weights = torch.full((5,), float(rank + 7))
    comm.broadcast_(weights)
    comm.barrier()
    np.savez(os.path.join(out_dir, f'rank{rank}.npz'), returns=returns, perms=np.stack(perms), gathered=gathered.numpy(),
             grads=grads.numpy(), slowest=slowest, decisions=np.array(decisions), weights=weights.numpy())
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_sharded_step_equals_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ranks = [np.load(tmp_path / f'rank{r}.npz') for r in range(world)]
    ro = synthetic.make_rollout(T, E, obs_shape=(4,), obs_dtype='float32', epochs=0, p_done=0.1)
    # returns: sharded == single-process, bit for bit (envs are independent)
    whole = oracle.gae_returns(ro.rewards, ro.dones, ro.values, ro.last_values, 0.99, 0.95)
    assert np.array_equal(np.concatenate([r['returns'] for r in ranks], axis=1), whole)
    # advantage statistics: combined parts == the single-process minibatch under the equivalent global permutation
    glob = xdist.global_minibatch_indices([list(r['perms']) for r in ranks], T, E, world, MB)
    flat_ret, flat_val = oracle.concat_step_batches(whole, ro.values)
    assert np.array_equal(ranks[0]['gathered'], ranks[1]['gathered'])
    for k in range(K):
        seen = np.concatenate(glob[k])
        assert len(seen) == len(set(seen.tolist())) == (T * E // (MB * world)) * MB * world
        for m in range(MB):
            adv = (flat_ret[glob[k][m]] - flat_val[glob[k][m]]).astype(np.float64)
            n, mean, std = xdist.combine_moments(ranks[0]['gathered'][:, k * MB + m])
            assert n == len(adv) and abs(mean - adv.mean()) < 1e-12 and abs(std - adv.std()) < 1e-12
            want = oracle.normalize_advantages(flat_ret[glob[k][m]], flat_val[glob[k][m]], 1e-8)
            got = ((adv - mean) / (std + 1e-8)).astype(np.float32)
            np.testing.assert_allclose(got, want, atol=1e-5 * np.abs(want).max())
    for r in ranks:
        assert np.array_equal(r['grads'], np.full(1000, 3.0)) and float(r['slowest']) == 20.0
        # job-wide mean reward (10 + 20 + 60) / 3 = 30 and step count 300: both ranks decide alike although rank 0 alone
        # (mean 15, 100 steps) would never have stopped and rank 1 alone (mean 60) would have stopped every time
        assert r['decisions'].tolist() == [True, False, True, False]
        assert np.array_equal(r['weights'], np.full(5, 7.0))
