"""CPU-side checks of the drop-in boundary: the library builds and loads, exports every symbol the
header declares, validates arguments before touching CUDA, and the DLPack reader is zero-copy.
No kernel runs here (there is no GPU in the build container)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from xagents_b200 import _build, _dlpack, _ffi, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    _build.build()
    return _ffi.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'xagents_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(xa_[a-z0-9_]+)\s*\(', text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 15
    out = subprocess.run(['nm', '-D', '--defined-only', _ffi.library_path()], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r' T (xa_\w+)', out))
    assert set(names) <= exported, sorted(set(names) - exported)
    assert set(names) == set(_ffi.PROTOTYPES), set(names) ^ set(_ffi.PROTOTYPES)
    assert lib.xa_version() == 10000


def test_library_is_sm100a_only_with_bulk_copy_sass():
    out = subprocess.run(['cuobjdump', '-lelf', _ffi.library_path()], capture_output=True, text=True).stdout
    assert 'sm_100a' in out and not re.search(r'sm_(?!100a)\d+', out), out
    sass = subprocess.run(['cuobjdump', '-sass', _ffi.library_path()], capture_output=True, text=True).stdout
    assert 'UBLKCP' in sass            # TMA bulk copies in the gather (B200_PROFILING.md evidence table)


def test_argument_errors_come_back_as_codes_not_crashes(lib):
    rc = lib.xa_gae_f32(None, None, None, None, None, None, 4, 3, 0.99, 0.95, 0, None)
    assert rc == -1 and b'null pointer' in lib.xa_last_error()
    rc = lib.xa_gae_f32(None, None, None, None, None, None, 0, 3, 0.99, 0.95, 0, None)
    assert rc == -1 and b'must be positive' in lib.xa_last_error()
    rc = lib.xa_gae_f32(8, 8, 8, 8, 8, None, 4, 3, 0.99, 0.95, 7, None)
    assert rc == -1 and b'unknown mode' in lib.xa_last_error()
    rc = lib.xa_gae_f32(8, 8, 8, 6, 8, None, 4, 3, 0.99, 0.95, 0, None)
    assert rc == -2 and b'aligned' in lib.xa_last_error()
    rc = lib.xa_gather_rows(None, None, None, 5, 16, 10, 0, 0, 0, None)
    assert rc == -1
    rc = lib.xa_gather_rows(16, 16, 16, 5, 16, 10, 3, 3, 0, None)       # T*E != n_src_rows
    assert rc == -1 and b'n_src_rows' in lib.xa_last_error()
    rc = lib.xa_gather_rows(16, 16, 16, 0, 16, 10, 0, 0, 0, None)       # empty gather is a no-op
    assert rc == 0
    args = _ffi.LossArgs()
    assert lib.xa_ppo_loss_f32(ctypes.byref(args), None) == -1
    assert lib.xa_ppo_loss_f32(None, None) == -1
    assert lib.xa_loss_workspace_bytes(8192) == 16 + 32 * 24
    assert lib.xa_clip_adam_f32(None, None, None, None, 10, None, 1e-3, .9, .999, 1e-7, 0.5, 1, 1.0, None) == -1
    with pytest.raises(_ffi.XAError) as e:
        _ffi.call('xa_nstep_returns_f32', None, None, None, None, 4, 3, 0.99, 0, None)
    assert e.value.code == -1 and 'null pointer' in str(e.value)


def test_no_cuda_device_is_an_error_not_a_fallback(lib):
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    assert lib.xa_device_info(0, None, None, None) != 0
    assert b'no CPU fallback' in lib.xa_last_error()


def test_loss_args_struct_matches_the_header_layout():
    # 64-bit: 7 pointers, 2 int32, 2 pointers, int32 (+pad), 2 int64, 2 int32, 4 floats, 5 pointers, int64
    assert ctypes.sizeof(_ffi.LossArgs) == 7 * 8 + 8 + 2 * 8 + 8 + 16 + 8 + 16 + 5 * 8 + 8
    assert _ffi.LossArgs.n.offset == 96 and _ffi.LossArgs.out_scalars.offset == 128


def test_network_plan_structs_match_the_header_layout(tmp_path):
    """xa_nature_cnn_t / xa_grad_segment_t as gcc lays them out from the header vs the ctypes mirrors."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / 'layout.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "xagents_b200.h"\nint main(void) {\n'
                   'printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(xa_grad_segment_t), sizeof(xa_nature_cnn_t), offsetof(xa_nature_cnn_t, w1),\n'
                   'offsetof(xa_nature_cnn_t, gemm_ws_bytes), offsetof(xa_nature_cnn_t, grad_map), offsetof(xa_nature_cnn_t, segments),\n'
                   'offsetof(xa_nature_cnn_t, n_grad)); return 0; }\n')
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-I', os.path.join(root, 'include'), str(src), '-o', str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    n = _ffi.NatureCnn
    assert got == [ctypes.sizeof(_ffi.GradSegment), ctypes.sizeof(n), n.w1.offset, n.gemm_ws_bytes.offset, n.grad_map.offset,
                   n.segments.offset, n.n_grad.offset]


def test_dlpack_reader_is_zero_copy_and_consumes_the_capsule():
    t = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4)
    arr = _dlpack.as_device_array(t, 'float32', allow_host=True)
    assert arr.ptr == t.data_ptr() and arr.shape == (2, 3, 4) and arr.dtype == 'float32' and arr.nbytes == 96
    view = _dlpack.host_numpy(arr)
    view[0, 0, 0] = 42.0                                  # same memory
    assert t[0, 0, 0].item() == 42.0
    arr.release()
    cap = t.__dlpack__()
    _dlpack.from_capsule(cap).release()
    with pytest.raises(TypeError):
        _dlpack.from_capsule(cap)                          # renamed to used_dltensor
    sl = torch.arange(10, dtype=torch.int32)[2:7]         # offset views keep their own pointer
    arr = _dlpack.as_device_array(sl, 'int32', allow_host=True)
    assert arr.ptr == sl.data_ptr() and arr.shape == (5,)
    with pytest.raises(ValueError):
        _dlpack.as_device_array(torch.zeros(4, 4).t(), allow_host=True)      # not C-contiguous
    with pytest.raises(TypeError):
        _dlpack.as_device_array(torch.zeros(4), 'int32', allow_host=True)
    with pytest.raises(TypeError):
        _dlpack.as_device_array([1, 2, 3])
    u8 = _dlpack.as_device_array(torch.zeros((2, 2), dtype=torch.uint8), allow_host=True)
    assert u8.dtype == 'uint8' and u8.itemsize == 1
    f64 = _dlpack.as_device_array(np.zeros(3), allow_host=True) if hasattr(np.zeros(3), '__dlpack__') else None
    assert f64 is None or f64.dtype == 'float64'


def test_ops_refuse_host_tensors():
    x = torch.zeros((4, 3))
    with pytest.raises(ValueError, match='no CPU path'):
        ops.gae_returns(x, x, torch.zeros(3), torch.zeros((5, 3)), 0.99, 0.95)
    with pytest.raises(ValueError, match='no CPU path'):
        ops.gather_rows(torch.zeros((4, 16), dtype=torch.uint8), torch.zeros(2, dtype=torch.int32))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setenv('XAGENTS_B200_LIB', str(tmp_path / 'nope.so'))
    monkeypatch.setattr(_ffi, '_lib', None)
    with pytest.raises(ImportError, match='no CPU or framework fallback'):
        _ffi.lib()


def test_bench_reference_arm_runs_the_stated_small_workloads_on_the_host():
    """`bench.py --impl reference --workload c1|c2`: the reference's CPU path (oracle port) on the host, one JSON line with the
    contract's keys; no GPU involved."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for workload, metric in (('c1', 'ppo_env_steps_per_sec_gae_gather_loss'), ('c2', 'a2c_env_steps_per_sec_returns_loss')):
        done = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--workload', workload, '--steps', '3',
                               '--warmup', '1'], capture_output=True, text=True, timeout=300, cwd=root)
        assert done.returncode == 0, done.stderr[-2000:]
        line = json.loads(done.stdout.strip().splitlines()[-1])
        assert line['impl'] == 'reference' and line['metric'] == metric and line['value'] > 0 and line['steps'] == 3
        assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] == 1 and workload in line['config']['workload']
        assert line['e2e']['h2d_bytes_per_step'] == 0 and line['gpu_launches'] == 0


def test_chained_launch_level_is_a_host_side_setting(lib):
    """xa_set_chained_launches: returns the previous level, ignores levels outside 0..3 (a query), needs no device."""
    before = lib.xa_set_chained_launches(-1)
    assert before in (0, 1, 2, 3)
    try:
        assert lib.xa_set_chained_launches(0) == before
        assert lib.xa_set_chained_launches(7) == 0 and lib.xa_set_chained_launches(-1) == 0      # out of range: nothing changes
        assert lib.xa_set_chained_launches(2) == 0 and lib.xa_set_chained_launches(-1) == 2
    finally:
        lib.xa_set_chained_launches(before)
    assert lib.xa_set_chained_launches(-1) == before


def test_new_entry_points_validate_before_touching_cuda(lib):
    splits = ctypes.c_int(5)
    rc = lib.xa_gemm_bf16_tn_partial(None, None, 256, 512, 3136, None, 0, ctypes.byref(splits), None)
    assert rc == -1 and b'null pointer' in lib.xa_last_error()
    rc = lib.xa_gemm_bf16_tn_partial(16, 16, 256, 512, 3133, None, 0, ctypes.byref(splits), None)
    assert rc == -1 and b'multiple of 8' in lib.xa_last_error()
    rc = lib.xa_heads_forward_partial_bf16(None, 2, None, None, None, None, None, None, 8, 512, 6, None)
    assert rc == -1 and b'null pointer' in lib.xa_last_error()
    rc = lib.xa_heads_forward_partial_bf16(16, 2, None, None, 16, 16, 16, 16, 8, 256, 6, None)
    assert rc == -1 and b'hidden' in lib.xa_last_error()
    rc = lib.xa_gemm_bf16_tn_maskbits(16, 16, 16, 64, 64, 64, 64, None, 64, 0, 0, None)
    assert rc == -1 and b'mask_bits' in lib.xa_last_error()
    rc = lib.xa_gemm_bf16_tn_maskbits(16, 16, 16, 64, 64, 64, 64, 16, 40, 0, 0, None)
    assert rc == -1 and b'multiple of 32' in lib.xa_last_error()
