"""One PPO update phase (run_ppo_epochs: 4 epochs x 4 minibatches of 8192 frames, network on the tcgen05 kernels) at C3's shape
inside cudaProfilerStart/Stop, for an ncu launch list:
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv python scripts/update_launches.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xagents_b200 import feeds
from xagents_b200.agents import PPO, NatureCnnTc, TorchModel

E, T, A = 256, 128, 6
dev = 'cuda:0'
torch.manual_seed(0)
envs = feeds.FedEnvs(E, (84, 84, 4), torch.uint8, A, device=dev)
net = TorchModel(NatureCnnTc(4, A).cuda())
agent = PPO(envs, net, n_steps=T, quiet=True, seed=1, device=dev)
if len(sys.argv) > 1 and sys.argv[1] == 'gather':
    agent.pipeline_options = dict(obs_gather=True)
gen = torch.Generator(device=dev)
gen.manual_seed(2)
agent.ro_states.copy_(torch.randint(0, 256, agent.ro_states.shape, dtype=torch.uint8, device=dev, generator=gen))
for name, scale in (('ro_rewards', 1.0), ('ro_values', 0.5), ('ro_log_probs', 0.3), ('ro_returns', 1.0)):
    getattr(agent, name).copy_(torch.randn(getattr(agent, name).shape, device=dev, generator=gen) * scale)
agent.ro_log_probs.abs_().neg_()
agent.ro_dones.copy_((torch.rand(agent.ro_dones.shape, device=dev, generator=gen) < 0.01).float())
agent.ro_actions.copy_(torch.randint(0, A, agent.ro_actions.shape, device=dev, generator=gen).float())
batch = agent.concat_step_batches(agent.ro_states, agent.ro_actions, agent.ro_returns, agent.ro_values, agent.ro_log_probs)
for _ in range(2):
    agent.run_ppo_epochs(*batch)
torch.cuda.synchronize()
import time
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
t0 = time.perf_counter()
agent.run_ppo_epochs(*batch)                                      # the queue is empty: the host is not throttled by the GPU
host_ms = (time.perf_counter() - t0) * 1e3
b.record()
torch.cuda.synchronize()
print('host issue time of one update phase (empty queue):', round(host_ms, 3), 'ms')
print('update phase, eager:', a.elapsed_time(b), 'ms')
torch.cuda.cudart().cudaProfilerStart()
agent.run_ppo_epochs(*batch)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
