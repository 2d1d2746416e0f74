"""One forward+backward of NatureCnnTc at B=8192 inside cudaProfilerStart/Stop, for an ncu launch list:
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv python scripts/cnn_launches.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xagents_b200.agents import NatureCnnTc

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
torch.manual_seed(0)
net = NatureCnnTc(4, 6).cuda().refresh()
frames = torch.randint(0, 256, (B, 84, 84, 4), dtype=torch.uint8, device='cuda')
da, dv = torch.randn(B, 6, device='cuda'), torch.randn(B, device='cuda')
for it in range(3):
    if it == 2:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    a, v = net(frames)
    torch.autograd.backward([a, v], [da, dv])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
