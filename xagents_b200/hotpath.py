"""Device-resident PPO / A2C rollout-to-update path for one rank's shard of environments.

`PPOHotPath` owns the time-major rollout buffers and the per-minibatch staging buffers and turns one
train step of the reference --

    PPO.get_batch:        calculate_returns (GAE)            xagents/ppo/agent.py:48-94, 193-213
                          concat_step_batches (flatten)      xagents/base.py:549-564      [fused away]
    PPO.run_ppo_epochs:   get_mini_batches (shuffle+gather)  xagents/ppo/agent.py:139-155
                          advantage normalisation            xagents/ppo/agent.py:180-183
    PPO.update_gradients: clipped surrogate+value+entropy    xagents/ppo/agent.py:96-134

-- into a handful of prepared launches through the C ABI: one GAE scan, one moments launch for all K*M minibatches, ONE
gather launch for all K*M minibatches (observation rows by TMA bulk copy; as few launches as fit the staging slots when
they do not fit one), and per minibatch a device-side wait for ITS rows plus one fused loss forward+backward.  Launch
arguments are resolved once (`prepare`), so a step is a tight loop of ctypes calls on fixed device pointers.

Two streams: a "data" stream carries the gather; the caller's stream carries the arithmetic (GAE -> moments -> [C2] ->
wait_0, forward_0, loss_0, backward_0 + C1 + optimiser_0, wait_1, ...).  The gather kernel counts the rows it has finished per
minibatch in device memory (`progress`), and the compute stream waits on the counter of the minibatch it is about to consume
(`sync='progress'`), so a long launch at full HBM rate feeds the per-minibatch chain without launch gaps; `sync='event'` (one
CUDA event per gather launch, several minibatches per launch on a tapered schedule) is the capturable alternative.  The loss
reads the four per-sample rollout scalars straight through the permutation (`fuse_fields=True`: no scalar gather, no
dependence of the gathers on GAE); `fuse_fields=False` materialises them like the reference does.

With a `dist.ShardComm` the environments are sharded across ranks (each rank owns a contiguous env-major range), the
per-minibatch advantage moments are all-gathered once per step (collective C2) so normalisation uses the GLOBAL minibatch
statistics as the single-process reference does; collective C1 belongs to the model adapter's `backward_and_step`
(`after_loss` hook).  The agents bind this pipeline in place to their own rollout buffers (`buffers=`, `bind`).
"""
import ctypes

import torch

from . import _ffi
from ._ffi import XA_MOMENT_STRIDE
from .ops import ACTOR_KINDS, GATHER_MODES, SCAN_MODES, _count


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(None)


class PPOHotPath:
    FIELDS = ('actions', 'returns', 'values', 'log_probs')     # gathered per minibatch, in this order

    def __init__(self, n_steps, n_envs, obs_shape, n_actions, *, obs_dtype=torch.uint8, ppo_epochs=4, mini_batches=4,
                 gamma=0.99, lam=0.95, clip_norm=0.1, entropy_coef=0.01, value_loss_coef=0.5, advantage_epsilon=1e-8,
                 actor_kind='logits', device='cuda:0', gather_mode='auto', scan_mode='auto', comm=None,
                 fuse_fields=True, staging=2, overlap=True, gather_chunk=None, buffers=None, sync='auto', dynamic=True, late_fork=False,
                 obs_gather=True):
        """`buffers`: existing time-major device tensors to run on instead of allocating (any of ROLLOUT_FIELDS and
        'returns') -- the agents pass their own `ro_*` rollout buffers, so the pipeline works in place (no copies).
        `obs_gather=False`: the consumer of the observations reads the rollout through `perms` itself (a network whose first
        layer fetches every frame by its id, TorchModel.reads_through_permutation): no frame gather is launched and no staging
        memory is allocated; `before_loss(i)` finds minibatch i's ids in `minibatch_ids(i)`."""
        self.obs_gather = bool(obs_gather)
        if not self.obs_gather:
            assert fuse_fields, 'obs_gather=False reads every field through the permutation'
            sync, gather_chunk, staging = 'event', int(ppo_epochs) * max(1, int(mini_batches)) * 2, 1
        self.T, self.E, self.A = int(n_steps), int(n_envs), int(n_actions)
        self.N = self.T * self.E
        obs_dtype = buffers['obs'].dtype if buffers and 'obs' in buffers else obs_dtype
        self.K, self.M = int(ppo_epochs), int(mini_batches)
        self.B = self.N // self.M                                  # mini_batch_size (ppo/agent.py:42-46)
        assert self.B > 0, f'Invalid batch size to mini-batch size ratio {self.N}: {self.M}'
        self.slices = [(lo, min(lo + self.B, self.N)) for lo in range(0, self.N, self.B)]   # ppo/agent.py:152
        self.n_mb = self.K * len(self.slices)
        self.obs_shape = tuple(obs_shape)
        self.gamma, self.lam = float(gamma), float(lam)
        self.clip_norm, self.entropy_coef = float(clip_norm), float(entropy_coef)
        self.value_loss_coef, self.advantage_epsilon = float(value_loss_coef), float(advantage_epsilon)
        self.actor_kind = actor_kind
        self.device = torch.device(device)
        if self.device.type == 'cuda' and self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        self.gather_mode, self.scan_mode = gather_mode, scan_mode
        self.comm = comm
        self.fuse_fields, self.staging, self.overlap = bool(fuse_fields), max(1, int(staging)), bool(overlap)
        # How the losses learn that their minibatch is staged.  'event': one CUDA event per gather launch (a loss waits for its
        # whole launch).  'progress': the gather kernel counts finished rows per minibatch in device memory and the compute
        # stream waits on the counter of ITS minibatch (a one-warp spin kernel; 'progress-memop': cuStreamWaitValue32, which
        # B200 re-polls only every ~3 ms once a wait has lasted ~0.5 ms), so one long launch -- full HBM rate, no launch
        # gaps -- feeds the per-minibatch chain loss -> backward -> all-reduce -> Adam as soon as each minibatch is complete.
        # 'auto' = 'progress' whenever the rows go through the TMA bulk path and the scalar fields are fused into the loss.
        assert sync in ('auto', 'event', 'progress', 'progress-memop'), f'unknown sync mode `{sync}`'
        # `dynamic` (progress mode): the gather's copy items are handed out through a device work counter instead of being
        # dealt statically, so the launch does not run at the pace of the CTAs that share an SM with the loss / optimiser
        # kernels.  `late_fork`: the data stream starts after GAE and the moments instead of beside them.
        self.dynamic, self.late_fork = bool(dynamic), bool(late_fork)
        row_b = self._row_bytes(obs_shape, obs_dtype)
        bulk_ok = gather_mode != 'vector' and row_b % 16 == 0 and self.fuse_fields and self.overlap and self.device.type == 'cuda'
        if sync.startswith('progress'):
            assert bulk_ok, 'progress sync needs 16-byte aligned rows (TMA bulk path), fused fields, the data stream and a CUDA device'
        worthwhile = row_b >= 2048 and self.B * row_b >= (16 << 20)          # where the gather's AUTO mode picks the bulk path
        self.sync = sync if sync.startswith('progress') else ('progress' if (sync == 'auto' and bulk_ok and worthwhile) else 'event')
        # minibatches moved per gather launch.  An int = fixed group size (1 = per minibatch, K*M = the whole
        # step, which is what get_mini_batches does: everything materialised before the first update); a list
        # = explicit schedule.  Default on one GPU: a taper -- half of what is left per launch, then (rest-1, 1):
        # [8, 4, 3, 1] for 16 minibatches.  Long launches saturate HBM (per-minibatch launches lose 7 % to
        # launch gaps at C3) and only ONE loss trails the last gather (measured 1.239 ms/step against 1.267 for
        # one launch per epoch).  With ranks > 1: per epoch, the last epoch per minibatch, so that a single
        # loss + gradient all-reduce trails (measured at 8 GPUs: four trailing all-reduces cost ~10 %).
        per_epoch = len(self.slices)
        if gather_chunk is None and self.sync.startswith('progress'):
            # as few launches as memory allows: <= 16 GB per staging slot; everything in one launch / one slot when it fits
            limit = max(1, int((16 << 30) // max(1, self.B * row_b)))
            sizes, rest = [], self.n_mb
            while rest > 0:
                sizes.append(min(rest, limit))
                rest -= sizes[-1]
        elif gather_chunk is None:
            if comm is not None and comm.world_size > 1:
                sizes = [per_epoch] * (self.K - 1) + [1] * per_epoch
            else:
                sizes, rest = [], self.n_mb
                limit = max(1, int((16 << 30) // max(1, self.B * self._row_bytes(obs_shape, obs_dtype))))   # <= 16 GB per slot
                while rest > 4:
                    take = min(rest // 2, limit)
                    sizes.append(take)
                    rest -= take
                sizes += [rest - 1, 1] if rest > 1 else [1]
        elif isinstance(gather_chunk, (list, tuple)):
            sizes = [int(x) for x in gather_chunk]
            assert sum(sizes) == self.n_mb and min(sizes) > 0, f'gather schedule {sizes} must cover {self.n_mb} minibatches'
        else:
            c = max(1, min(int(gather_chunk), self.n_mb))
            sizes = [c] * (self.n_mb // c) + ([self.n_mb % c] if self.n_mb % c else [])
        self.group_sizes = sizes
        self.group_first = [sum(sizes[:g]) for g in range(len(sizes))]
        self.chunk = max(sizes)
        self.n_groups = len(sizes)
        self.staging = min(self.staging, self.n_groups)
        mb_rows = [hi - lo for _ in range(self.K) for lo, hi in self.slices]
        self.group_rows = [sum(mb_rows[f:f + n]) for f, n in zip(self.group_first, sizes)]
        self.mb_rows = mb_rows
        self.cap = max(self.group_rows)
        dev, f32 = self.device, torch.float32
        T, E, N, B, A = self.T, self.E, self.N, self.B, self.A
        if comm is not None and comm.world_size > 1:
            # equal shards: the loss divides by the LOCAL minibatch size and the gradients are averaged over ranks (a mean
            # of per-rank means is the global mean only then), and every rank must issue the same number of collectives
            assert comm.max_over_ranks(N) == N and comm.max_over_ranks(-N) == -N, (
                f'sharded PPO needs the same n_steps * n_envs on every rank (this rank: {N})')
        # rollout, time-major (what the rollout loop writes step by step)
        shapes = {'obs': ((T, E) + self.obs_shape, obs_dtype), 'rewards': ((T, E), f32), 'values': ((T, E), f32),
                  'last_values': ((E,), f32), 'dones': ((T + 1, E), f32), 'actions': ((T, E), f32), 'log_probs': ((T, E), f32),
                  'returns': ((T, E), f32)}
        for name, (shape, dtype) in shapes.items():
            setattr(self, name, torch.empty(shape, dtype=dtype, device=dev) if not (buffers and name in buffers) else None)
        if buffers:
            self.bind(**buffers)
        # permutations of env-major flat sample ids, one row per epoch
        self.perms = torch.empty((self.K, N), dtype=torch.int32, device=dev)
        # minibatch staging
        self.mb_obs = torch.empty((self.staging, self.cap) + self.obs_shape, dtype=obs_dtype, device=dev) if self.obs_gather else None
        self.mb_fields = torch.empty((self.staging, len(self.FIELDS), self.cap), dtype=f32, device=dev)
        # model outputs for every minibatch (filled by the caller / the model forward)
        self.actor_out = torch.empty((self.n_mb, B, A), dtype=f32, device=dev)
        self.critic_out = torch.empty((self.n_mb, B), dtype=f32, device=dev)
        # results
        self.moments = torch.zeros((self.n_mb, XA_MOMENT_STRIDE), dtype=torch.float64, device=dev)
        world = comm.world_size if comm is not None else 1
        self.all_moments = (torch.zeros((world, self.n_mb, XA_MOMENT_STRIDE), dtype=torch.float64, device=dev)
                            if world > 1 else None)
        self.scalars = torch.zeros((self.n_mb, 4), dtype=f32, device=dev)
        self.d_actor = torch.empty((B, A), dtype=f32, device=dev)
        self.d_values = torch.empty((B,), dtype=f32, device=dev)
        nbytes = _ffi.lib().xa_loss_workspace_bytes(B)
        self.workspace = torch.zeros((nbytes + 7) // 8, dtype=torch.float64, device=dev)
        self._calls = None
        self.on_gather = None          # default `on_gather` hook of run() (bench.py times the gather launches through it)
        self.row_bytes = self.obs.element_size()
        for k in self.obs_shape:
            self.row_bytes *= k

    @staticmethod
    def _row_bytes(obs_shape, obs_dtype):
        n = torch.empty((), dtype=obs_dtype).element_size()
        for k in obs_shape:
            n *= int(k)
        return n

    # ---------------------------------------------------------------------------------- data in
    def bind(self, **buffers):
        """Run on the caller's rollout tensors (time-major, contiguous, on this device).  Continuous actions may be
        [T, E, k] (read through the permutation by the loss; they cannot ride along as a gathered scalar field)."""
        T, E = self.T, self.E
        for name, t in buffers.items():
            assert name in self.ROLLOUT_FIELDS + ('returns',), f'unknown rollout field `{name}`'
            assert t.is_contiguous() and t.device == self.device, f'{name}: need a contiguous tensor on {self.device}'
            lead = (E,) if name == 'last_values' else ((T + 1, E) if name == 'dones' else (T, E))
            assert tuple(t.shape[:len(lead)]) == lead, f'{name}: shape {tuple(t.shape)} is not time-major {lead}'
            if name == 'obs':
                assert tuple(t.shape[2:]) == self.obs_shape, f'obs rows {tuple(t.shape[2:])} != {self.obs_shape}'
            elif name == 'actions':
                assert t.dim() == 2 or self.fuse_fields, 'vector actions need fuse_fields=True'
                assert t.dtype == torch.float32
            else:
                assert t.dtype == torch.float32 and t.dim() == len(lead), f'{name}: need fp32 {lead}'
            setattr(self, name, t)
        self._calls = None
        return self

    ROLLOUT_FIELDS = ('obs', 'rewards', 'values', 'last_values', 'dones', 'actions', 'log_probs')

    def load(self, rollout, stream=None, non_blocking=True):
        """Copy a rollout (host tensors / numpy / synthetic.Rollout) into the device buffers."""
        with torch.cuda.stream(stream) if stream is not None else _null():
            for name in self.ROLLOUT_FIELDS:
                src = getattr(rollout, name) if not isinstance(rollout, dict) else rollout[name]
                getattr(self, name).copy_(torch.as_tensor(src), non_blocking=non_blocking)

    def h2d_bytes(self):
        return sum(getattr(self, n).numel() * getattr(self, n).element_size() for n in self.ROLLOUT_FIELDS)

    # ---------------------------------------------------------------------------------- plan
    def prepare(self, stream=None):
        """Resolve every launch of one train step to (cfunc, args) on fixed device pointers."""
        lib = _ffi.lib()
        if self.device.type != 'cuda':
            # planning only (host-logic tests): the launch tables are built, but there is no CPU path -- run()
            # fails in the first launch
            self.compute_stream = self.data_stream = None
            sc = sd = ctypes.c_void_p(None)
        else:
            self.compute_stream = stream if stream is not None else torch.cuda.current_stream(self.device)
            self.data_stream = torch.cuda.Stream(self.device) if self.overlap else self.compute_stream
            sc = ctypes.c_void_p(self.compute_stream.cuda_stream)
            sd = ctypes.c_void_p(self.data_stream.cuda_stream)
        T, E, N, B = self.T, self.E, self.N, self.B
        self._gae = (lib.xa_gae_f32, (_p(self.rewards), _p(self.values), _p(self.last_values), _p(self.dones),
                                      _p(self.returns), _p(None), T, E, self.gamma, self.lam, SCAN_MODES[self.scan_mode], sc))
        offsets = []
        for k in range(self.K):
            offsets += [k * N + lo for lo, _ in self.slices]
        offsets.append(self.K * N)
        # the last slice of epoch k ends where epoch k+1 starts, so one ascending table serves all K*M
        self._offsets = (ctypes.c_int64 * len(offsets))(*offsets)
        self._moments = (lib.xa_adv_moments_f32, (_p(self.returns), _p(self.values), _p(self.perms), self._offsets, self.n_mb,
                                                  T, E, _p(self.moments), sc))
        world = self.comm.world_size if self.comm is not None else 1
        n_fields = 0 if self.fuse_fields else len(self.FIELDS)
        self._field_src = (ctypes.c_void_p * 4)(*[getattr(self, f).data_ptr() for f in self.FIELDS])
        self._field_dst, self._loss_args, self._gathers, self._losses = [], [], [], []
        fsz = 4 * self.cap
        for slot in range(self.staging):
            base = self.mb_fields.data_ptr() + slot * len(self.FIELDS) * fsz
            self._field_dst.append((ctypes.c_void_p * 4)(*[base + j * fsz for j in range(4)]))
        flat_off = list(self._offsets)                       # start of every minibatch in perms.view(-1)
        self._mb_place = []                                  # minibatch -> (group, slot, first row in the slot)
        self._progress_sync = self.sync.startswith('progress') and self.device.type == 'cuda' and not getattr(self, '_capturing', False)
        if self._progress_sync:
            if getattr(self, 'progress', None) is None:
                self.progress = torch.zeros(self.n_mb, dtype=torch.int32, device=self.device)      # cyclic counters, never reset
                self._step_no = 0
            self._units = lib.xa_gather_progress_units(self.row_bytes)
            self._wait_memop = self.sync == 'progress-memop'
            self._wait = lib.xa_stream_wait_geq_u32 if self._wait_memop else lib.xa_wait_progress_u32
            if getattr(self, 'wait_status', None) is None:
                self.wait_status = torch.zeros(1, dtype=torch.int32, device=self.device)     # 1 = a wait gave up after ~2 s
            self._wait_status = _p(self.wait_status)
            self._wait_addr = [ctypes.c_void_p(self.progress.data_ptr() + 4 * i) for i in range(self.n_mb)]
            self._wait_stream = sc
            if self.dynamic and getattr(self, 'work', None) is None:
                self.work = torch.zeros(2 * self.n_groups, dtype=torch.int32, device=self.device)   # per launch: next item, lanes done
        for g in range(self.n_groups):
            first, slot = self.group_first[g], g % self.staging
            idx_addr = self.perms.data_ptr() + 4 * flat_off[first]
            obs_dst = ctypes.c_void_p(self.mb_obs.data_ptr() + slot * self.cap * self.row_bytes) if self.obs_gather else None
            if not self.obs_gather:
                self._gathers.append(None)
            elif self._progress_sync:
                self._gathers.append((lib.xa_gather_rows_progress,
                                      (_p(self.obs), ctypes.c_void_p(idx_addr), obs_dst, self.group_rows[g], self.row_bytes, N, T, E,
                                       _p(self.progress), flat_off[first], N, B,
                                       ctypes.c_void_p(self.work.data_ptr() + 8 * g) if self.dynamic else ctypes.c_void_p(None), sd)))
            else:
                self._gathers.append((lib.xa_gather_minibatch,
                                      (_p(self.obs), obs_dst, self.row_bytes, N, self._field_src, self._field_dst[slot],
                                       n_fields, ctypes.c_void_p(idx_addr), self.group_rows[g], T, E,
                                       GATHER_MODES[self.gather_mode], sd)))
            for mb in range(first, first + self.group_sizes[g]):
                self._mb_place.append((g, slot, flat_off[mb] - flat_off[first]))
        for mb in range(self.n_mb):
            g, slot, row0 = self._mb_place[mb]
            n = self.mb_rows[mb]
            a = _ffi.LossArgs()
            a.actor_out = self.actor_out.data_ptr() + 4 * mb * B * self.A
            a.values = self.critic_out.data_ptr() + 4 * mb * B
            if self.fuse_fields:         # read the time-major rollout through the permutation
                a.actions, a.returns, a.old_values, a.old_log_probs = (getattr(self, f).data_ptr() for f in self.FIELDS)
                a.idx, a.n_steps, a.n_envs = self.perms.data_ptr() + 4 * flat_off[mb], T, E
            else:
                base = self.mb_fields.data_ptr() + slot * len(self.FIELDS) * fsz + 4 * row0
                a.actions, a.returns, a.old_values, a.old_log_probs = (base + j * fsz for j in range(4))
                a.idx, a.n_steps, a.n_envs = None, 0, 0
            a.advantages = None
            if world > 1:
                a.moments = self.all_moments.data_ptr() + 8 * mb * XA_MOMENT_STRIDE
                a.n_moment_parts, a.moment_part_stride = world, self.n_mb * XA_MOMENT_STRIDE
            else:
                a.moments = self.moments.data_ptr() + 8 * mb * XA_MOMENT_STRIDE
                a.n_moment_parts, a.moment_part_stride = 1, 0
            a.n, a.n_actions, a.actor_kind = n, self.A, ACTOR_KINDS[self.actor_kind]
            a.clip, a.ent_coef, a.vf_coef, a.adv_eps = (self.clip_norm, self.entropy_coef, self.value_loss_coef,
                                                        self.advantage_epsilon)
            a.out_scalars = self.scalars.data_ptr() + 16 * mb
            a.d_actor, a.d_values, a.advantages_out = self.d_actor.data_ptr(), self.d_values.data_ptr(), None
            a.workspace, a.workspace_bytes = self.workspace.data_ptr(), self.workspace.numel() * 8
            self._loss_args.append(a)
            self._losses.append((lib.xa_ppo_loss_f32, (ctypes.byref(a), sc)))
        if self.device.type == 'cuda':
            self._gather_done = [torch.cuda.Event() for _ in range(self.n_groups)]
            self._loss_done = [torch.cuda.Event() for _ in range(self.n_groups)]
            self._fork = torch.cuda.Event()
        self._calls = True
        self.kernel_launches_per_step = 2 + self.n_mb + (self.n_groups if self.obs_gather else 0)
        return self

    # ---------------------------------------------------------------------------------- run
    def _check(self, tag, rc):
        if rc != 0:
            raise _ffi.XAError(tag, rc, _ffi.lib().xa_last_error().decode('utf-8', 'replace'))

    def run(self, on_gather=None, after_loss=None, before_loss=None, gae=True):
        """Issue one train step.  `on_gather(i, fn, args)` may wrap gather i (bench timing, on
        `self.data_stream`); `before_loss(i)` runs on the compute stream once minibatch i is staged and must
        leave the model outputs in `actor_out[i]` / `critic_out[i]` (model forward); `after_loss(i)` runs after
        minibatch i's loss was enqueued (model backward from `d_actor` / `d_values`, collective C1, optimiser).
        `gae=False`: `returns` already holds this rollout's returns (the agents' `calculate_returns` wrote them)."""
        if self._calls is None:
            self.prepare()
        if self.device.type == 'cuda' and torch.cuda.current_device() != self.device.index:
            with torch.cuda.device(self.device):                    # launches go to the CUDA runtime's current device
                return self.run(on_gather, after_loss, before_loss, gae)
        cs, ds = self.compute_stream, self.data_stream
        two = ds is not cs
        on_gather = on_gather if on_gather is not None else self.on_gather
        # the data stream joins after everything already queued on the compute stream (rollout writes,
        # the previous step) -- and after GAE only when the gathers carry the returns field
        early = two and self.fuse_fields and not self.late_fork
        if early:
            self._fork.record(cs)
            ds.wait_event(self._fork)
        _count(self.kernel_launches_per_step - (0 if gae else 1))
        if gae:
            fn, args = self._gae
            self._check('gae', fn(*args))
        if two and not self.fuse_fields:
            self._fork.record(cs)
            ds.wait_event(self._fork)
        fn, args = self._moments
        self._check('moments', fn(*args))
        if two and self.fuse_fields and self.late_fork:
            self._fork.record(cs)
            ds.wait_event(self._fork)
        if self.comm is not None and self.comm.world_size > 1:
            with torch.cuda.stream(cs):                                         # ordered between the moments and the losses
                self.comm.all_gather_moments(self.all_moments, self.moments)    # collective C2
        progress = self._progress_sync
        if progress:
            self._step_no += 1
        for g in range(self.n_groups):
            if self._gathers[g] is not None:
                fn, args = self._gathers[g]
                if two and g >= self.staging:
                    ds.wait_event(self._loss_done[g - self.staging])            # staging slot is free again
                self._check('gather', on_gather(g, fn, args) if on_gather is not None else fn(*args))
                if two:
                    self._gather_done[g].record(ds)
                    if not progress:
                        cs.wait_event(self._gather_done[g])
            for i in range(self.group_first[g], self.group_first[g] + self.group_sizes[g]):
                if progress:     # minibatch i is staged once its counter reached (launches so far) x (its rows) x (units per row)
                    target = (self._step_no * self.mb_rows[i] * self._units) & 0xffffffff
                    if self._wait_memop:
                        self._check('wait', self._wait(self._wait_stream, self._wait_addr[i], target))
                    else:
                        self._check('wait', self._wait(self._wait_addr[i], target, self._wait_status, self._wait_stream))
                if before_loss is not None:
                    before_loss(i)
                fn, args = self._losses[i]
                self._check('loss', fn(*args))
                if after_loss is not None:
                    after_loss(i)
            if two:
                self._loss_done[g].record(cs)
        if progress:
            cs.wait_event(self._gather_done[self.n_groups - 1])                 # formal join: the gather KERNELS are done too

    # ---------------------------------------------------------------------------------- CUDA graph
    def capture(self):
        """Capture one train step (both streams, every launch) into a CUDA graph and return a replay callable.
        For latency-bound shapes (CartPole-sized rollouts move a few hundred KB) the step is nothing but launch
        overhead; a replay is one launch.  Buffers keep their addresses, so new rollouts / permutations / model
        outputs are simply written into them before each replay."""
        if self._calls is None:
            self.prepare()
        assert self.comm is None or self.comm.world_size == 1, 'graph capture covers the single-GPU pipeline'
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        capture_stream = torch.cuda.Stream(self.device)
        eager_streams = (self.compute_stream, self.data_stream)
        self._capturing = True                               # a captured step orders by events (no stream memory ops in the graph)
        with torch.cuda.stream(capture_stream):
            self.prepare(capture_stream)                    # re-resolve the launches onto the capturing stream
            self.run()                                      # warm-up outside the capture (lazy module loading)
            capture_stream.synchronize()
            with torch.cuda.graph(graph, stream=capture_stream):
                self.run()
        self._graph = graph                                  # keep alive
        self._capturing = False
        self.prepare(eager_streams[0])                       # eager path stays usable
        return graph.replay

    def minibatch_ids(self, i):
        """int32 [rows of minibatch i]: its env-major sample ids, a view of `perms` (what every kernel reads it through)."""
        return self.perms.view(-1)[self._offsets[i]:self._offsets[i] + self.mb_rows[i]]

    def minibatch_views(self, i):
        """(states, actions, returns, old_values, old_log_probs) staging views of minibatch i."""
        _, slot, row0 = self._mb_place[i]
        n = self.mb_rows[i]
        f = self.mb_fields[slot]
        return (self.mb_obs[slot, row0:row0 + n],) + tuple(f[j, row0:row0 + n] for j in range(len(self.FIELDS)))

    # ---------------------------------------------------------------------------------- accounting
    def algorithmic_bytes(self):
        """SURVEY.md 8d / BASELINE.md 3: bytes one train step must move (read + write, no re-reads)."""
        N, E, K, A, F = self.N, self.E, self.K, self.A, self.row_bytes
        gae = 16 * N + 4 * E
        gather = K * (2 * F + 36) * N
        moments = K * 8 * N
        loss = K * (8 * A + 24) * N
        rows = sum(self.group_rows) / len(self.group_rows)
        per_launch = (2 * F + 4) * rows if self.fuse_fields else (2 * F + 36) * rows
        return dict(gae=gae, gather=gather, moments=moments, loss=loss, total=gae + gather + moments + loss,
                    gather_per_launch=per_launch)


class _null:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


class A2CHotPath:
    """A2C.train_step's arithmetic on the device (xagents/a2c/agent.py:173-218): n-step returns over the
    time-major rollout, then the loss + gradients over all T*E samples in time-major order (the loss is a mean
    over the batch, so the env-major flatten of np_train_step changes nothing but the summation order).
    Two launches per train step."""

    FIELDS = ('rewards', 'values', 'last_values', 'dones', 'actions', 'returns')

    def __init__(self, n_steps, n_envs, n_actions, *, gamma=0.99, entropy_coef=0.01, value_loss_coef=0.5, actor_kind='logits',
                 device='cuda:0', scan_mode='auto', buffers=None):
        """`buffers`: the caller's time-major rollout tensors (any of FIELDS) to work on in place, as in PPOHotPath."""
        self.T, self.E, self.A = int(n_steps), int(n_envs), int(n_actions)
        self.N = self.T * self.E
        self.gamma, self.entropy_coef, self.value_loss_coef = float(gamma), float(entropy_coef), float(value_loss_coef)
        self.actor_kind, self.scan_mode = actor_kind, scan_mode
        self.device = torch.device(device)
        if self.device.type == 'cuda' and self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        dev, f32, T, E, N, A = self.device, torch.float32, self.T, self.E, self.N, self.A
        shapes = {'rewards': (T, E), 'values': (T, E), 'last_values': (E,), 'dones': (T + 1, E), 'actions': (T, E), 'returns': (T, E)}
        for name, shape in shapes.items():
            setattr(self, name, torch.empty(shape, dtype=f32, device=dev) if not (buffers and name in buffers) else None)
        self.actor_out = torch.empty((N, A), dtype=f32, device=dev)       # forward pass outputs, time-major sample order
        self.critic_out = torch.empty((N,), dtype=f32, device=dev)
        self.scalars = torch.zeros(4, dtype=f32, device=dev)
        self.d_actor = torch.empty((N, A), dtype=f32, device=dev)
        self.d_values = torch.empty((N,), dtype=f32, device=dev)
        nbytes = _ffi.lib().xa_loss_workspace_bytes(N)
        self.workspace = torch.zeros((nbytes + 7) // 8, dtype=torch.float64, device=dev)
        self.kernel_launches_per_step = 2
        self._args = None
        if buffers:
            self.bind(**buffers)

    def bind(self, **buffers):
        for name, t in buffers.items():
            assert name in self.FIELDS, f'unknown rollout field `{name}`'
            assert t.is_contiguous() and t.device == self.device and t.dtype == torch.float32, (
                f'{name}: need a contiguous fp32 tensor on {self.device}')
            setattr(self, name, t)
        self._args = None
        return self

    def prepare(self, stream=None):
        lib = _ffi.lib()
        self.stream = stream if stream is not None else torch.cuda.current_stream(self.device)
        s = ctypes.c_void_p(self.stream.cuda_stream)
        self._returns = (lib.xa_nstep_returns_f32, (_p(self.rewards), _p(self.dones), _p(self.last_values), _p(self.returns), self.T,
                                                    self.E, self.gamma, SCAN_MODES[self.scan_mode], s))
        a = _ffi.LossArgs()
        a.actor_out, a.values = self.actor_out.data_ptr(), self.critic_out.data_ptr()
        a.actions, a.old_values, a.returns = self.actions.data_ptr(), self.values.data_ptr(), self.returns.data_ptr()
        a.old_log_probs = a.idx = a.advantages = a.moments = None
        a.n, a.n_actions, a.actor_kind = self.N, self.A, ACTOR_KINDS[self.actor_kind]
        a.ent_coef, a.vf_coef = self.entropy_coef, self.value_loss_coef
        a.out_scalars, a.d_actor, a.d_values = self.scalars.data_ptr(), self.d_actor.data_ptr(), self.d_values.data_ptr()
        a.workspace, a.workspace_bytes = self.workspace.data_ptr(), self.workspace.numel() * 8
        self._args = a
        self._loss = (lib.xa_a2c_loss_f32, (ctypes.byref(a), s))
        return self

    def run(self, returns=True, loss=True):
        """`returns=False`: `self.returns` is already filled (the agent's `calculate_returns` wrote it); `loss=False`: only the
        returns scan (the model forward, which the loss needs, comes after it in A2C.train_step)."""
        if self._args is None:
            self.prepare()
        if self.device.type == 'cuda' and torch.cuda.current_device() != self.device.index:
            with torch.cuda.device(self.device):
                return self.run(returns, loss)
        for tag, on, (fn, args) in (('nstep_returns', returns, self._returns), ('a2c_loss', loss, self._loss)):
            if not on:
                continue
            _count()
            rc = fn(*args)
            if rc != 0:
                raise _ffi.XAError(tag, rc, _ffi.lib().xa_last_error().decode('utf-8', 'replace'))
