import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xagents_b200 import ops
B, H, W, C, kh, kw, N = [int(v) for v in os.environ.get('SHAPE', '5,9,8,64,2,1,64').split(',')]
g = torch.Generator(device='cuda'); g.manual_seed(1)
x = torch.randn((B, H, W, C), device='cuda', generator=g).to(torch.bfloat16)
OH, OW = H - kh + 1, W - kw + 1
dy = torch.randn((B, OH, OW, N), device='cuda', generator=g).to(torch.bfloat16)
want = torch.nn.grad.conv2d_weight(x.double().permute(0, 3, 1, 2), (N, C, kh, kw), dy.double().permute(0, 3, 1, 2)).permute(0, 2, 3, 1).reshape(N, -1)
dw, db = ops.conv_wgrad_bf16(dy.reshape(-1, N), x, kh, kw)
torch.cuda.synchronize()
print('shape', (B, H, W, C, kh, kw, N), 'max err / scale', float((dw.double() - want).abs().max() / want.abs().max()))
