import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xagents_b200 import _ffi, ops
lib = _ffi.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr())
def timeit(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
B = 8192
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for name, (H, W, C, kh, kw, N, s2d) in {'conv3': (9, 9, 64, 3, 3, 64, 0), 'conv2': (10, 10, 128, 2, 2, 64, 0), 'conv1': (21, 21, 64, 2, 2, 32, 1)}.items():
    x = torch.randn((B, H, W, C), device='cuda').to(torch.bfloat16)
    OH, OW = H - kh + 1, W - kw + 1
    dy = torch.randn((B * OH * OW, N), device='cuda').to(torch.bfloat16)
    Wg = -(-W // 8) * 8
    Q = B * H * Wg
    dyt = torch.empty((kw, N, Q), dtype=torch.bfloat16, device='cuda')
    xt = torch.empty((C, Q), dtype=torch.bfloat16, device='cuda')
    ws = torch.empty(lib.xa_conv_wgrad_workspace_bytes(N, C, kh, kw) // 4, dtype=torch.float32, device='cuda')
    dw = torch.empty((N, kh * kw * C), dtype=torch.float32, device='cuda')
    t1 = timeit(lambda: lib.xa_place_on_grid_t_bf16(P(dy), P(dyt), N, B, H, Wg, OH, OW, Q, s2d, kw, st))
    t2 = timeit(lambda: lib.xa_place_on_grid_t_bf16(P(x), P(xt), C, B, H, Wg, H, W, Q, 0, 1, st))
    t3 = timeit(lambda: lib.xa_conv_wgrad_bf16(P(dyt), P(xt), P(dw), N, C, kh, kw, Wg, Q, Q, P(ws), ws.numel() * 4, st))
    t4 = timeit(lambda: dyt[0].sum(1, dtype=torch.float32))
    print(f'{name}: place dY ({kw} copies, {dyt.numel()*2/1e6:.0f} MB) {t1:.0f} us | place X ({xt.numel()*2/1e6:.0f} MB) {t2:.0f} us | wgrad kernels {t3:.0f} us | bias sum {t4:.0f} us')
