"""A few environment steps of the PPO rollout at C3's shape (eager loop: the same kernels the captured graph replays) inside
cudaProfilerStart/Stop, for an ncu launch list:
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv python scripts/rollout_launches.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xagents_b200 import envs as xenvs
from xagents_b200.agents import PPO, NatureCnnTc, TorchModel

E, T, A = 256, 4, 6
torch.manual_seed(0)
made = xenvs.create_envs('SyntheticAtariDevice-v0', E, preprocess=True, device='cuda:0')
net = TorchModel(NatureCnnTc(4, A).cuda())
agent = PPO(made, net, n_steps=T, mini_batches=4, ppo_epochs=1, quiet=True, seed=1)
agent.graph_rollout = False
agent.fit(max_steps=2 * T * E)                # warm-up
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
agent.get_batch()                             # T environment steps + the bootstrap forward + GAE
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
