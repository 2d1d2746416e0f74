#!/usr/bin/env python
"""`xagents train ...` under torchrun with the environments sharded over the ranks, plus the check that makes the run
evidence: after training, every rank holds bit-identical network weights (same initial broadcast, same averaged
gradients, same fused clip+Adam), took the same number of optimiser steps, and the job-wide step count is what the
command line asked for.  Rank 0 prints one JSON line.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/sharded_train_check.py train ppo --env CartPole-v1 --n-envs 16 --max-steps 40960 --seed 1 --quiet
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from xagents_b200 import cli  # noqa: E402


def main():
    ex = cli.Executor()
    t0 = time.perf_counter()
    ex.execute(sys.argv[1:])
    elapsed = time.perf_counter() - t0
    agent = ex.agent
    world = dist.get_world_size() if dist.is_initialized() else 1
    net = agent.net
    torch.cuda.synchronize()
    equal, steps, rewards = True, agent.steps, [float(sum(agent.total_rewards)), float(len(agent.total_rewards))]
    if world > 1:
        parts = [torch.empty_like(net.flat_param) for _ in range(world)]
        dist.all_gather(parts, net.flat_param)
        equal = all(torch.equal(parts[0], p) for p in parts[1:])
        t = torch.tensor([agent.steps, net.step] + rewards, dtype=torch.float64, device=agent.device)
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        steps = int(sum(g[0].item() for g in gathered))
        assert len({int(g[1].item()) for g in gathered}) == 1, 'ranks took different numbers of optimiser steps'
        rewards = [sum(g[2].item() for g in gathered), sum(g[3].item() for g in gathered)]
    if (int(os.environ.get('RANK', '0'))) == 0:
        print(json.dumps({'world_size': world, 'agent': type(agent).__name__, 'n_envs_per_rank': agent.n_envs, 'job_steps': steps,
                          'optimizer_steps': net.step, 'weights_identical_across_ranks': bool(equal),
                          'weights_finite': bool(torch.isfinite(net.flat_param).all()),
                          'mean_reward_last_episodes': rewards[0] / rewards[1] if rewards[1] else None,
                          'seconds': round(elapsed, 2)}))
    if world > 1:
        dist.barrier(device_ids=[torch.device(agent.device).index])
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
