"""Environment sharding across the GPUs of one box and the two collectives the path needs.

The reference is single-process (its "multiple environments" are a Python loop, xagents/base.py:402-425);
sharding is new here.  Environments are independent units: rank g of G owns envs [g*E/G, (g+1)*E/G).
Because the flat sample layout is env-major (b = e*T + t, xagents/base.py:559-564) a rank's samples are
one contiguous flat range, so returns/GAE and the gathers need NO communication.  Two collectives:

  C2  one all-gather per train step of the per-minibatch advantage moments (count, mean, M2) so that
      the normalisation of xagents/ppo/agent.py:180-183 uses the statistics of the GLOBAL minibatch;
  C1  a sum all-reduce of the flat gradient buffer per minibatch, then 1/G, global-norm clip and Adam.

On a box whose GPUs can map each other's memory both run as small kernels of this library over NVLink peer memory
(xagents_b200/peer.py, csrc/peer_adam.cu: C1 fused with the optimiser; C2 a 128-thread kernel), because NCCL's kernels
cannot share an SM with the persistent gather and would run after it.  This module keeps the NCCL versions (the fallback and
the measured baseline: all-reduce on a high-priority side stream, waited for by the optimiser) and the host-side plumbing.

One process per GPU, torch.distributed for the plumbing (NCCL over NVLink on the box, gloo in CPU tests).
"""
import os

import torch
import torch.distributed as dist


def shard_range(n_envs, rank, world_size):
    """[lo, hi) of the envs rank owns; shards differ by at most one env when G does not divide E."""
    base, extra = divmod(n_envs, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_time_major(array, rank, world_size):
    """Slice a time-major [T, E, ...] (or [E]) rollout field down to this rank's envs."""
    env_axis = 0 if array.ndim == 1 else 1
    lo, hi = shard_range(array.shape[env_axis], rank, world_size)
    return array[lo:hi] if env_axis == 0 else array[:, lo:hi]


def global_minibatch_indices(local_perms, n_steps, n_envs, world_size, mini_batches):
    """The single-process permutation that is equivalent to the sharded run.

    local_perms[g][k] is rank g's epoch-k permutation of its local flat ids.  Global minibatch (k, m) is
    the rank-wise concatenation of the local minibatches (k, m), local id b on rank g mapping to global
    flat id (env_lo_g * T + b).  Returns [K][M] lists of global env-major ids (numpy int64).
    """
    import numpy as np
    K = len(local_perms[0])
    out = []
    for k in range(K):
        epoch = []
        for m in range(mini_batches):
            parts = []
            for g in range(world_size):
                lo, hi = shard_range(n_envs, g, world_size)
                n_local = (hi - lo) * n_steps
                b_local = n_local // mini_batches
                sl = np.asarray(local_perms[g][k][m * b_local:(m + 1) * b_local], dtype=np.int64)
                parts.append(sl + lo * n_steps)
            epoch.append(np.concatenate(parts))
        out.append(epoch)
    return out


class ShardComm:
    """Thin wrapper over a torch.distributed process group for the path's two collectives."""

    def __init__(self, group=None, device=None):
        assert dist.is_initialized(), 'call init_from_env() (or dist.init_process_group) first'
        self.group = group
        self.rank = dist.get_rank(group)
        self.world_size = dist.get_world_size(group)
        self.backend = dist.get_backend(group)
        self.device = device
        # High priority: measured on 8xB200, four 6.75 MB all-reduces issued under a 574 us gather cost 310 us
        # of exposed time on a normal-priority stream (the NCCL CTAs queue behind the gather's) and 103 us
        # on a high-priority one.
        self.comm_stream = (torch.cuda.Stream(device, priority=-1)
                            if (device is not None and torch.device(device).type == 'cuda') else None)
        self._pending = None
        self.peer_c2 = 'auto'            # 'auto': C2 through NVLink peer memory when the box can map it (peer.PeerAllGather); False: NCCL
        self._peer_gathers = {}

    # C2 ------------------------------------------------------------------------------------------
    def all_gather_moments(self, out, local):
        """out [G, n_mb, 4] <- every rank's local [n_mb, 4] moments (fp64), on the current stream."""
        if self.peer_c2 and self.backend == 'nccl' and local.is_cuda:
            # NCCL's kernels cannot share an SM with the persistent gather's CTAs (register file) and would run after it -- with
            # the whole per-minibatch chain behind them; a 128-thread P2P kernel does not have that problem
            from . import peer
            if peer.available(self):
                key = local.numel()
                if key not in self._peer_gathers:
                    self._peer_gathers[key] = peer.PeerAllGather(self, key)
                self._peer_gathers[key].gather(local, out)
                return out
            self.peer_c2 = False
        if self.backend == 'gloo':
            parts = [torch.empty_like(local) for _ in range(self.world_size)]
            dist.all_gather(parts, local, group=self.group)
            out.copy_(torch.stack(parts))
        else:
            dist.all_gather_into_tensor(out.view(-1), local.view(-1), group=self.group)
        return out

    # C1 ------------------------------------------------------------------------------------------
    def all_reduce_gradients_async(self, flat_grads):
        """Sum-all-reduce the flat gradient buffer on the side stream (after what is already queued on
        the current stream), leaving the current stream free to start the next minibatch's gather."""
        if self.comm_stream is None:
            dist.all_reduce(flat_grads, group=self.group)
            return
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        self.comm_stream.wait_event(ready)
        with torch.cuda.stream(self.comm_stream):
            dist.all_reduce(flat_grads, group=self.group)
            self._pending = torch.cuda.Event()
            self._pending.record(self.comm_stream)

    def wait_gradients(self):
        """Make the current stream wait for the last all-reduce (call before the optimiser step)."""
        if self._pending is not None:
            torch.cuda.current_stream(self.device).wait_event(self._pending)
            self._pending = None

    def sum_over_ranks(self, values):
        """Element-wise sum of a short list of host numbers over all ranks (stop conditions, logging)."""
        t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=self.device if self.backend == 'nccl' else 'cpu')
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.tolist()

    def broadcast_(self, tensor, src=0):
        dist.broadcast(tensor, src, group=self.group)
        return tensor

    def max_over_ranks(self, value):
        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device if self.backend == 'nccl' else 'cpu')
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def barrier(self):
        if self.backend == 'nccl':
            dist.barrier(group=self.group, device_ids=[torch.device(self.device).index])
        else:
            dist.barrier(group=self.group)


def init_from_env(backend=None):
    """Join the job torchrun described (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*). Returns (rank, local_rank, world)."""
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
        # The path's collectives are latency-bound (6.75 MB gradients, a few hundred bytes of moments): on
        # 8xB200 the ring/tree kernels measured faster than NVLS here (60 vs 66 us alone, and they overlap
        # the gathers better).  Respect an explicit user setting.
        os.environ.setdefault('NCCL_NVLS_ENABLE', '0')
        kw = {}
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
            kw['device_id'] = torch.device('cuda', local_rank)
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def combine_moments(parts):
    """Host mirror of the kernels' Chan combination: parts [G, 4] = (count, mean, M2, _) -> (count, mean, std)."""
    cn = cm = c2 = 0.0
    for qn, qm, q2, *_ in (tuple(float(x) for x in p) for p in parts):
        if qn <= 0.0:
            continue
        tot, delta = cn + qn, qm - cm
        c2 += q2 + delta * delta * cn * qn / tot
        cm += delta * qn / tot
        cn = tot
    return cn, cm, (c2 / cn) ** 0.5 if cn > 0 else 0.0
