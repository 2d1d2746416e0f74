"""Build the C-ABI shared library in-tree with nvcc for sm_100a.

    python -m xagents_b200._build          # -> xagents_b200/lib/libxagents_b200.so

nvcc cross-compiles without a GPU; the runtime is linked statically so the library depends only on
the driver (libcuda) of the box it runs on.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
LIB_DIR = os.path.join(PKG, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libxagents_b200.so')
SOURCES = ['capi.cu', 'returns_scan.cu', 'gather.cu', 'loss.cu', 'optim.cu', 'peer_adam.cu', 'policy.cu', 'synth_env.cu', 'gemm_tc.cu', 'conv_tc.cu', 'conv_flat_tc.cu', 'wgrad_mn_tc.cu', 'gemm_atb_tc.cu', 'heads.cu', 'grad_finalize.cu', 'nature_net.cu']
ARCH_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a']


def nvcc_path():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found (set NVCC=/path/to/nvcc)')


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, 'include', 'xagents_b200.h')]
    return any(os.path.getmtime(d) > built for d in deps)


def build(force=False, verbose=False):
    """Compile every kernel for sm_100a into one shared library. Returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [nvcc_path(), *ARCH_FLAGS, '--threads', '0', '-lineinfo', '-O3', '-std=c++17', '-shared', '-Xcompiler', '-fPIC',
           '-I', os.path.join(ROOT, 'include'), '-o', LIB_PATH + '.tmp', *srcs]
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
        print(' '.join(cmd))
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f'nvcc failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}')
    if verbose:
        print(proc.stderr)
    os.replace(LIB_PATH + '.tmp', LIB_PATH)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
