"""Collective C1 fused with the optimiser over NVLink peer memory (csrc/peer_adam.cu).

`FusedAllReduceAdam` owns the peer-mapped ("symmetric") buffers of one rank -- the flat gradient, the flat parameter
vector, G partial sums and the barrier flags -- plus the rank's SHARD of Adam's m and v, and replaces

    ncclAllReduce(flat_grad); xa_grad_sumsq_f32; xa_clip_adam_f32          (3 launches, NCCL's CTAs beside the gather's)

by ONE cooperative kernel per update: reduce-scatter by P2P loads, global norm, clip + Adam on the owned shard,
all-gather of the new weights by P2P stores (SURVEY.md 8e C1 + 8f-2).  The peer mappings come from
`torch.distributed._symmetric_memory` (cuMem allocations exchanged over the process group's store): plumbing only --
the raw pointers go to the C ABI (`xa_peer_allreduce_adam_f32`), no torch collective runs in a step.

There is no fallback inside this class: if the box cannot map peer memory, `available()` says so and the callers keep
the NCCL path (dist.ShardComm.all_reduce_gradients_async + ops.clip_adam).
"""
import ctypes

import torch
import torch.distributed as dist

from . import _ffi

_PROBE = {}


def _symm():
    import torch.distributed._symmetric_memory as symm_mem
    return symm_mem


class _RawDeviceBuffer:
    """A cudaMalloc'ed region as a torch tensor (zero-copy, through __cuda_array_interface__)."""

    def __init__(self, ptr, words):
        self.__cuda_array_interface__ = {'shape': (words,), 'typestr': '<f4', 'data': (int(ptr), False), 'version': 2}


def _map_symm(comm, words):
    """-> (local fp32 tensor [words], [base pointer of every rank's buffer in this process], keep-alive)."""
    symm_mem = _symm()
    buf = symm_mem.empty(words, dtype=torch.float32, device=comm.device)
    buf.zero_()
    hdl = symm_mem.rendezvous(buf, comm.group if comm.group is not None else dist.group.WORLD)
    ptrs = [int(p) for p in hdl.buffer_ptrs]
    assert len(ptrs) == comm.world_size and all(ptrs), f'symmetric-memory rendezvous returned {ptrs}'
    return buf, ptrs, hdl


def _map_ipc(comm, words):
    lib = _ffi.lib()
    ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
    with torch.cuda.device(comm.device):
        _ffi.check('xa_ipc_alloc', lib.xa_ipc_alloc(4 * words, ctypes.byref(ptr), handle))
        handles = [None] * comm.world_size
        dist.all_gather_object(handles, bytes(handle.raw), group=comm.group)
        ptrs = []
        for r, h in enumerate(handles):
            if r == comm.rank:
                ptrs.append(int(ptr.value))
                continue
            remote = ctypes.c_void_p()
            _ffi.check('xa_ipc_open', lib.xa_ipc_open(h, ctypes.byref(remote)))
            ptrs.append(int(remote.value))
    buf = torch.as_tensor(_RawDeviceBuffer(ptr.value, words), device=comm.device)
    return buf, ptrs, ('ipc', ptr.value)      # lives for the rest of the process (IPC mappings are not torn down mid-job)


_TRANSPORTS = {'symm': _map_symm, 'ipc': _map_ipc}


def transport(comm):
    """Which peer-memory transport works on this box for `comm`: 'symm' (torch symmetric memory), 'ipc' (CUDA IPC
    handles exchanged over the process group) or None.  One collective probe per process, cached; same answer on all ranks."""
    if comm is None or comm.world_size < 2 or comm.world_size > _ffi.XA_MAX_PEERS or comm.backend != 'nccl':
        return None
    if 'transport' not in _PROBE:
        _PROBE['transport'], _PROBE['errors'] = None, {}
        for name in ('symm', 'ipc'):
            ok = 1.0
            try:
                buf, ptrs, keep = _TRANSPORTS[name](comm, 1024)
                _PROBE.setdefault('keep', []).append((buf, keep))
            except Exception as exc:   # noqa: BLE001 -- any failure means "this transport is not available here"
                _PROBE['errors'][name] = f'{type(exc).__name__}: {exc}'
                ok = 0.0
            if comm.max_over_ranks(-ok) == -1.0:        # min over ranks
                _PROBE['transport'] = name
                break
    return _PROBE['transport']


def available(comm):
    return transport(comm) is not None


def probe_errors():
    return dict(_PROBE.get('errors', {}))


class FusedAllReduceAdam:
    """One rank's side of the fused gradient all-reduce + global-norm clip + Adam."""

    def __init__(self, comm, n_params, *, lr=7e-4, beta1=0.9, beta2=0.999, eps=1e-7, init=None, via=None):
        G, rank, dev = comm.world_size, comm.rank, torch.device(comm.device)
        assert 2 <= G <= _ffi.XA_MAX_PEERS, f'fused C1 supports 2..{_ffi.XA_MAX_PEERS} ranks of one box'
        via = via or transport(comm)
        assert via in _TRANSPORTS, f'no peer-memory transport on this box ({probe_errors()}): keep the NCCL path'
        self.comm, self.device, self.lr, self.beta1, self.beta2, self.eps, self.via = comm, dev, lr, beta1, beta2, eps, via
        per = -(-int(n_params) // G)
        self.shard_len = per + (-per) % 4
        self.n_params, self.n_pad = int(n_params), self.shard_len * G
        lib = _ffi.lib()
        flag_words = lib.xa_peer_adam_flag_bytes() // 4
        # one peer-mapped allocation: [grad n_pad | param n_pad | sumsq 2*XA_MAX_PEERS words | flags] as 32-bit words
        words = 2 * self.n_pad + 2 * _ffi.XA_MAX_PEERS + flag_words
        self._buf, ptrs, self._keep = _TRANSPORTS[via](comm, words)
        self.grad = self._buf[:self.n_pad]
        self.param = self._buf[self.n_pad:2 * self.n_pad]
        if init is not None:
            self.param[:self.n_params].copy_(init.reshape(-1)[:self.n_params])
        self.m = torch.zeros(self.shard_len, dtype=torch.float32, device=dev)
        self.v = torch.zeros(self.shard_len, dtype=torch.float32, device=dev)
        self._ws = torch.zeros((lib.xa_peer_adam_workspace_bytes() + 7) // 8, dtype=torch.float64, device=dev)
        a = self._args = _ffi.PeerAdamArgs()
        for r, base in enumerate(ptrs):
            a.grad[r] = base
            a.param[r] = base + 4 * self.n_pad
            a.sumsq[r] = base + 8 * self.n_pad
            a.flags[r] = base + 8 * self.n_pad + 8 * _ffi.XA_MAX_PEERS
        a.m, a.v, a.workspace = self.m.data_ptr(), self.v.data_ptr(), self._ws.data_ptr()
        a.shard_len, a.rank, a.world = self.shard_len, rank, G
        self.epoch = 0
        torch.cuda.synchronize(dev)
        comm.barrier()                    # every rank's buffers are zeroed / initialised before anybody's first step

    def step(self, step, grad_norm=None, stream=None):
        """Average `self.grad` over the ranks, clip by the global norm, Adam-update `self.param` everywhere.  Asynchronous
        on `stream` (default: torch's current stream); every rank must call it the same number of times."""
        self.epoch += 1
        self._args.epoch = self.epoch
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        _ffi.call('xa_peer_allreduce_adam_f32', ctypes.byref(self._args), float(self.lr), float(self.beta1), float(self.beta2),
                  float(self.eps), float(grad_norm) if grad_norm else 0.0, int(step), ctypes.c_void_p(s.cuda_stream))

    def status(self):
        """0 = every wait so far completed; 1..3 = the phase whose wait timed out (synchronises)."""
        return int(self._ws.view(torch.int32)[2].item())


class PeerAllGather:
    """Collective C2 without NCCL: every rank's small fp64 block lands in every rank's table through P2P stores (one 128-thread
    kernel per call, csrc/peer_adam.cu).  `gather(local, out)`: out [world, ...] <- every rank's `local`, on the current stream."""

    def __init__(self, comm, n_per_rank, via=None):
        G, dev = comm.world_size, torch.device(comm.device)
        via = via or transport(comm)
        assert via in _TRANSPORTS, f'no peer-memory transport on this box ({probe_errors()})'
        self.comm, self.device, self.n = comm, dev, int(n_per_rank)
        words = 2 * (2 * G * self.n) + _ffi.XA_MAX_PEERS            # two generations of [G, n] doubles, then the flags
        self._buf, ptrs, self._keep = _TRANSPORTS[via](comm, words)
        self._status = torch.zeros(1, dtype=torch.int32, device=dev)
        a = self._args = _ffi.PeerGatherArgs()
        for r, base in enumerate(ptrs):
            a.recv[r] = base
            a.flags[r] = base + 8 * (2 * G * self.n)
        a.status, a.n_per_rank, a.rank, a.world = self._status.data_ptr(), self.n, comm.rank, G
        self.epoch = 0
        torch.cuda.synchronize(dev)
        comm.barrier()

    def gather(self, local, out, stream=None):
        assert local.dtype == torch.float64 and out.dtype == torch.float64 and local.numel() == self.n
        assert out.numel() == self.n * self.comm.world_size and local.is_contiguous() and out.is_contiguous()
        self.epoch += 1
        self._args.epoch = self.epoch
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        _ffi.call('xa_peer_allgather_f64', ctypes.byref(self._args), ctypes.c_void_p(local.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                  ctypes.c_void_p(s.cuda_stream))
        return out

    def status(self):
        return int(self._status.item())
