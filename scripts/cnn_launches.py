"""One forward+backward of the tensor-core network at B=8192 inside cudaProfilerStart/Stop, for an ncu launch list:
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv python scripts/cnn_launches.py [B] [autograd]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xagents_b200.agents import NatureCnnTc, TorchModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
autograd = len(sys.argv) > 2 and sys.argv[2] == 'autograd'
torch.manual_seed(0)
tm = TorchModel(NatureCnnTc(4, 6).cuda(), native_plan=not autograd)
tm.lr = 0.0
frames = torch.randint(0, 256, (B, 84, 84, 4), dtype=torch.uint8, device='cuda')
da, dv = torch.randn(B, 6, device='cuda') / B, torch.randn(B, device='cuda') / B
for it in range(3):
    if it == 2:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    tm.forward(frames, training=True)
    if autograd:
        tm.flat_grad.zero_()
        torch.autograd.backward(list(tm._outputs), [da, dv])
    else:
        tm._outputs.backward(da, dv, tm.flat_grad)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
