"""The first layer straight from uint8 frames (xa_conv2d_u8_s2d_bf16) alone, for ncu: python scripts/conv1_u8_probe.py [B]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xagents_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
torch.manual_seed(0)
x = torch.randint(0, 256, (B, 84, 84, 4), dtype=torch.uint8, device='cuda')
w = (torch.randn(32, 256, device='cuda') * 0.05).bfloat16()
b = torch.zeros(32, device='cuda')
y = torch.empty((B, 10, 10, 128), dtype=torch.bfloat16, device='cuda')
for _ in range(3):
    ops.conv2d_u8_s2d_bf16(x, w, 2, 2, bias=b, relu=True, out_s2d=True, out=y)
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    ops.conv2d_u8_s2d_bf16(x, w, 2, 2, bias=b, relu=True, out_s2d=True, out=y)
e.record()
torch.cuda.synchronize()
print(f'B={B}: {a.elapsed_time(e) / 20 * 1e3:.1f} us')
