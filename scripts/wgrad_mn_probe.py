"""Correctness + timing probe for xa_conv_wgrad_nhwc_bf16 (XA_WGRAD_MODE = 0 | 1 | 2 picks the operand addressing)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xagents_b200 import ops

torch.manual_seed(0)
dev = 'cuda'


def reference(x, dyg, kh, kw):
    B, H, W, C = x.shape
    N = dyg.shape[-1]
    xf, df = x.double(), dyg.double()
    dw = torch.zeros(N, kh, kw, C, dtype=torch.float64, device=x.device)
    for i in range(kh):
        for j in range(kw):
            # dYg is zero where the shift would leave the image, so the flat shift equals the 2-D shift
            dw[:, i, j] = torch.einsum('byxn,byxc->nc', df[:, :H - i, :W - j], xf[:, i:, j:])
    return dw.reshape(N, -1), df.sum((0, 1, 2))


def case(B, H, W, C, N, kh, kw, timing=False):
    OH, OW = H - kh + 1, W - kw + 1
    x = torch.randn(B, H, W, C, device=dev).to(torch.bfloat16)
    dyg = torch.zeros(B, H, W, N, device=dev, dtype=torch.bfloat16)
    dyg[:, :OH, :OW] = torch.randn(B, OH, OW, N, device=dev).to(torch.bfloat16)
    dw, db = ops.conv_wgrad_nhwc_bf16(x, dyg, kh, kw)
    torch.cuda.synchronize()
    rw, rb = reference(x, dyg, kh, kw)
    ew = ((dw.double() - rw).norm() / rw.norm()).item()
    eb = ((db.double() - rb).norm() / rb.norm()).item()
    msg = f'B={B} grid {H}x{W} C={C} N={N} k={kh}x{kw}: rel err dW {ew:.2e} db {eb:.2e}'
    if timing:
        for _ in range(3):
            ops.conv_wgrad_nhwc_bf16(x, dyg, kh, kw)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            ops.conv_wgrad_nhwc_bf16(x, dyg, kh, kw)
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1000 / 20
        gb = (x.numel() + dyg.numel()) * 2 / 1e9
        msg += f' | {us:.0f} us, {gb / us * 1e6:.0f} GB/s of operand bytes'
    print(msg, flush=True)
    return ew, eb


if len(sys.argv) > 1 and sys.argv[1] == 'ncu':       # one launch per layer shape, nothing else: for an ncu capture
    for B, H, W, C, N, kh, kw in [(8192, 9, 9, 64, 64, 3, 3), (8192, 10, 10, 128, 64, 2, 2), (8192, 21, 21, 64, 32, 2, 2)]:
        x = torch.randn(B, H, W, C, device=dev).to(torch.bfloat16)
        dyg = torch.randn(B, H, W, N, device=dev).to(torch.bfloat16)
        ops.conv_wgrad_nhwc_bf16(x, dyg, kh, kw)
    torch.cuda.synchronize()
    sys.exit(0)
print('mode', os.environ.get('XA_WGRAD_MODE', '0'))
big = len(sys.argv) > 1 and sys.argv[1] == 'big'
for shp in [(3, 9, 9, 64, 64, 3, 3), (5, 10, 10, 128, 64, 2, 2), (2, 21, 21, 64, 32, 2, 2), (37, 9, 9, 64, 64, 3, 3)]:
    case(*shp)
if big:
    for shp in [(8192, 9, 9, 64, 64, 3, 3), (8192, 10, 10, 128, 64, 2, 2), (8192, 21, 21, 64, 32, 2, 2)]:
        case(*shp, timing=True)
