#!/usr/bin/env python
"""Headline benchmark: PPO env-steps/s through GAE + permute-gather + loss (BASELINE.json metric).

    python bench.py [--gpus N --steps K --warmup W]            our arm (one process per GPU under torchrun)
    python bench.py --impl reference [--gpus N ...]            the reference's CPU path on the host cores
    python bench.py --workload c1|c2|c5                        the small / sweep configs of BASELINE.json (secondary lines)

A "step" is one `PPO.train_step()` of the drop-in agent (xagents_b200.agents.PPO) over one synthetic rollout that a feed
leaves in the agent's time-major buffers: 1 GAE scan, the advantage moments of all K*M minibatches, then K*M x
(minibatch gather of uint8 frames, fused loss forward+backward reading the rollout scalars through the permutation,
gradient stand-in -> [all-reduce] -> fused global-norm clip + Adam on the Nature CNN's 1.69 M parameters).  The
policy/value network's contractions are not part of the metric: its outputs are inputs, and a stand-in
kernel turns the loss's output gradients into the flat parameter gradient so that collective C1 and the optimiser sit
where they sit in training -- all-reduce i after backward i, Adam i after all-reduce i, loss i+1 after Adam i.
N=1 runs config C3 (n_envs=256, n_steps=128, 4 epochs x 4 minibatches, 84x84x4 uint8 frames); N>1 runs C4
(n_envs=4096 sharded over the ranks).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'ppo_env_steps_per_sec_gae_gather_loss'
UNIT = 'env-steps/s'
NATURE_CNN_PARAMS = 1_687_719          # Conv2D Nature CNN @84x84x4, 6 actions: C1 payload
CARTPOLE_MLP_PARAMS = 4_675            # the default CartPole .cfg network (tests/test_gpu_agents.py)
FALLBACK_HBM_GBS = 6650.0              # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
FRAME = (84, 84, 4)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default=None, choices=[None, 'c1', 'c2', 'c3', 'c4', 'c5'])
    ap.add_argument('--n-envs', type=int, default=None, help='total n_envs (overrides the workload default)')
    ap.add_argument('--n-steps', type=int, default=None)
    ap.add_argument('--gather-mode', default='auto', choices=['auto', 'bulk', 'vector'])
    ap.add_argument('--no-overlap', action='store_true', help='gathers on the compute stream instead of a data stream')
    ap.add_argument('--staging', type=int, default=2)
    ap.add_argument('--static-gather', action='store_true', help='deal the gather items statically instead of through a work counter')
    ap.add_argument('--late-fork', action='store_true', help='start the data stream after GAE + moments instead of beside them')
    ap.add_argument('--obs-gather', default='auto', choices=['auto', 'on', 'off'],
                    help='off: the network reads the rollout through the permutation (no frame gather); auto: off where the network can (nature-tc)')
    ap.add_argument('--sync', default='auto', choices=['auto', 'event', 'progress', 'progress-memop'], help='how losses learn their minibatch is staged (hotpath.PPOHotPath)')
    ap.add_argument('--gather-chunk', type=int, default=None, help='minibatches per gather launch (default: tapered schedule)')
    ap.add_argument('--gather-schedule', default=None, help='explicit launch schedule, e.g. 4,4,4,3,1')
    ap.add_argument('--c1', default='auto', choices=['auto', 'nccl', 'fused', 'none'],
                    help='gradient all-reduce per minibatch: NCCL all-reduce + clip/Adam, the fused peer-memory kernel, or none (diagnostic)')
    ap.add_argument('--no-optimizer', action='store_true', help='diagnostic: no gradient stand-in / all-reduce / Adam after the loss')
    ap.add_argument('--network', default='stand-in', choices=['stand-in', 'nature-tc'],
                    help='nature-tc: the real Nature CNN forward/backward on the tcgen05 kernels inside the step (secondary metric)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--time-updates', action='store_true', help='diagnostic: CUDA events around every gradient stand-in + C1 + optimiser update')
    ap.add_argument('--no-clock-sampler', action='store_true', help='diagnostic: no NVML polling thread beside the timed region')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-single-gpu-compare', action='store_true', help='N>1: skip the 1-GPU runs at the same per-GPU n_envs / at C4')
    ap.add_argument('--no-parity-check', action='store_true', help='N>1: skip the sharded-vs-oracle check before timing')
    ap.add_argument('--cpu-sample-envs', type=int, default=None, help='reference arm: time an n_envs sample (stated in config.workload)')
    ap.add_argument('--out', default=None, help='c5: also write the sweep as markdown to this path')
    return ap.parse_args()


def hbm_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return FALLBACK_HBM_GBS, 'fallback (B200_PROFILING.md)'


def workload_of(args, world):
    """-> (name, total n_envs, n_steps, obs shape, obs dtype name, n_actions, description)."""
    name = args.workload or ('c3' if world == 1 else 'c4')
    if name == 'c1':
        E, T, shape, dtype, A = 16, 128, (4,), 'float32', 2
        what = 'PPO on CartPole-shaped rollouts (fp32 [4] observations, 2 actions)'
    elif name == 'c2':
        E, T, shape, dtype, A = 16, 5, FRAME, 'uint8', 6
        what = 'A2C on synthetic Pong-shaped frames (84x84x4 uint8)'
    else:
        E, T, shape, dtype, A = (256 if name == 'c3' else 4096), 128, FRAME, 'uint8', 6
        what = 'PPO on synthetic Atari frames (84x84x4 uint8)'
    E, T = args.n_envs or E, args.n_steps or T
    desc = f'{name}: {what}, n_envs={E}{" sharded over %d ranks" % world if world > 1 else ""}, n_steps={T}'
    if name != 'c2':
        desc += ', 4 epochs x 4 minibatches'
    return name, E, T, shape, dtype, A, desc + f', {A} actions'


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock, power and throttle reasons sampled WHILE the bench runs.  NVML in a thread (a query takes well under
    a millisecond, so a 25 ms timed region still gets a dozen samples); `nvidia-smi -lms` if NVML cannot be loaded."""
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, gpu_index, period_s=0.002):
        self.gpu_index, self.period_s = gpu_index, period_s
        self.proc, self.thread, self.lines, self.samples, self.stop_flag, self.source = None, None, [], [], False, None

    def _nvml_loop(self, nv, handle, masks):
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
                mem = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_MEM)
                power = nv.nvmlDeviceGetPowerUsage(handle) / 1000.0
                bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                self.samples.append((float(sm), power, [n for n, m in masks if bits & m], float(mem)))
            except Exception:
                pass
            time.sleep(self.period_s)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            handle = nv.nvmlDeviceGetHandleByIndex(self.gpu_index)
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
            masks = [('hw_slowdown', nv.nvmlClocksThrottleReasonHwSlowdown),
                     ('hw_thermal_slowdown', nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                     ('sw_thermal_slowdown', nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                     ('sw_power_cap', nv.nvmlClocksThrottleReasonSwPowerCap)]
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle, masks), daemon=True)
            self.thread.start()
            self.source = 'nvml'
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.gpu_index}', f'--query-gpu={self.QUERY}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            self.source = 'nvidia-smi'
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Samples taken from here on are the ones reported (call right before the timed region)."""
        self.first = len(self.samples)

    def stop(self):
        if self.source == 'nvml':
            self.stop_flag = True
            self.thread.join(timeout=1)
            first = getattr(self, 'first', 0)
            used = self.samples[first:] or self.samples[-1:]
            sm = [x[0] for x in used]
            reasons = sorted({r for x in used for r in x[2]})
            mem = [x[3] for x in used]
            return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': self.sm_max, 'sm_min_mhz_seen': min(sm) if sm else None,
                    'mem_mhz': statistics.median(mem) if mem else None, 'mem_min_mhz_seen': min(mem) if mem else None,
                    'power_w_max': max((x[1] for x in used), default=None),
                    'power_w_median': statistics.median([x[1] for x in used]) if used else None, 'samples': len(sm), 'reasons': reasons,
                    'source': 'nvml, sampled every %.0f ms from the start of the timed region' % (self.period_s * 1e3)}
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, flag in zip(self.NAMES, parts[4:8]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(power) if power else None, 'samples': len(sm), 'reasons': sorted(reasons),
                'source': 'nvidia-smi -lms 100'}


# ------------------------------------------------------------------------------------------------ reference arm
def host_threads():
    # all the host threads the box has: torchrun exports OMP_NUM_THREADS=1, which would make the N>1 reference arm slower
    # than the N=1 one for no reason of the reference's
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference(name, n_envs, n_steps, obs_shape, obs_dtype, n_actions, steps, warmup):
    """The reference's CPU path (oracle port; TF is not installable here) on the host cores.  -> (result, cpu_baseline dict)"""
    import torch
    # latency-bound shapes (C1: 2048 samples of 16 B, C2: 80 frames): torch's intra-op thread pool only adds hand-off latency
    # to ops on a few KB (measured here: the 80-sample A2C loss takes 100 ms on 8 threads, 1 ms on one) -- one thread is the
    # faster CPU configuration, so that is what is reported
    torch.set_num_threads(1 if name in ('c1', 'c2') else host_threads())
    from oracle import cpu_path
    from xagents_b200 import synthetic
    ro = synthetic.make_rollout(n_steps, n_envs, obs_shape=obs_shape, obs_dtype=obs_dtype, n_actions=n_actions,
                                epochs=0 if name == 'c2' else 4)
    if name == 'c2':
        res = cpu_path.time_cpu_baseline(ro, steps=steps, warmup=warmup, algo='a2c', layout='reference')
        res_u8 = cpu_path.time_cpu_baseline(ro, steps=max(1, steps // 2), warmup=1, algo='a2c', layout='uint8')
    else:
        res = cpu_path.time_cpu_baseline(ro, steps=steps, warmup=warmup, layout='reference')
        # also with the observations kept uint8 on the host (a quarter of the bytes the reference's fp32 layout
        # moves) -- the fairest CPU number for the byte movement itself; `value` stays the reference's own layout
        res_u8 = cpu_path.time_cpu_baseline(ro, steps=max(1, steps // 2), warmup=1, layout='uint8')
    sample = (f'{steps} timed train steps (median; after {warmup} warm-up) of n_envs={n_envs} x n_steps={n_steps} '
              f'(= {n_envs * n_steps} samples/step), observations fp32 as the reference stores them')
    base = {'value': res['env_steps_per_sec'], 'unit': UNIT, 'cores': res['threads'], 'kind': 'port', 'sample': sample,
            'value_uint8_obs': res_u8['env_steps_per_sec'], 'host_cpus': os.cpu_count(), 'torch_threads': torch.get_num_threads(),
            'note': 'oracle port: NumPy for what the reference does in NumPy, torch-CPU ops for its TF-CPU ops (TensorFlow is not '
                    'installable in this image); gradients / tfp arithmetic follow the published definitions, not a TF run'}
    return res, base


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if rank != 0:
        return
    name, E, T, shape, dtype, A, desc = workload_of(args, max(world, args.gpus))
    if name == 'c5':
        emit({'impl': 'reference', 'unavailable': 'c5 is a kernel sweep: its CPU column is inside the `--workload c5` line'})
        return
    steps, warmup = args.steps, args.warmup
    sample_envs = args.cpu_sample_envs
    if sample_envs is None and name == 'c4':
        # E=4096 stores 59 GB of fp32 observations on the host and takes ~30 s per train step: the arm times an n_envs=256
        # sample of it (4096/256 = 16 of these per step; the CPU rate does not improve with size: it is memory-bound)
        sample_envs = 256
    if sample_envs is not None and sample_envs != E:
        desc += f' -- CPU ARM TIMED ON AN n_envs={sample_envs} SAMPLE of this workload'
        E = sample_envs
    if name in ('c3', 'c4'):            # ~1.8 s per train step at E=256 on 16 cores: keep the whole run within a few minutes
        budget_s = 150.0
        est = 1.8 * E / 256
        if (steps + warmup) * est > budget_s:
            warmup = min(warmup, 2)
            steps = max(3, int(budget_s / est) - warmup)
            desc += f' ({steps} timed steps after {warmup} warm-up: bounded to ~{budget_s:.0f} s)'
    res, base = cpu_reference(name, E, T, shape, dtype, A, steps, warmup)
    line = {
        'impl': 'reference', 'metric': METRIC if name != 'c2' else 'a2c_env_steps_per_sec_returns_loss', 'value': res['env_steps_per_sec'],
        'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup, 'ms_per_step': res['seconds_per_step'] * 1e3,
        'higher_is_better': True, 'scaling': 'weak' if args.gpus == 1 else 'strong', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': desc, 'note': 'reference CPU path (NumPy + torch-CPU stand-ins for the TF-CPU ops; '
                                             'TensorFlow is not installable in this image), host cores only'},
        'cpu_baseline': base,
        'e2e': {'value': res['env_steps_per_sec'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ our arm
class SyntheticNetwork:
    """What the agent needs from "the model" (agents/models.py protocol) when the network's contractions are outside the
    metric: its outputs per minibatch are inputs (precomputed tables), and its backward is the stand-in kernel that turns
    the loss's output gradients into the flat parameter gradient.  Collective C1 and the fused clip+Adam are the real ones,
    on Nature-CNN-sized buffers."""
    output_is_softmax = False

    def __init__(self, device, comm, n_params, optimizer=True, c1='nccl', lr=7e-4):
        import ctypes

        import torch

        from xagents_b200 import _ffi, ops
        from xagents_b200.optim import FlatAdam
        self.torch, self.ops, self.comm, self.optimizer, self.c1, self.lr = torch, ops, comm, optimizer, c1, lr
        self.device = torch.device(device)
        n = n_params + (-n_params) % 4
        self.n_params = n_params
        self.step = 0
        self.tables = None               # (actor [n_mb, B, A], critic [n_mb, B]) uploaded per rollout (e2e); None: already in place
        self.fused = self.opt = None
        if c1 == 'fused' and comm is not None and comm.world_size > 1:
            from xagents_b200 import peer
            self.fused = peer.FusedAllReduceAdam(comm, n_params, lr=lr)
            self.flat_grad, self.flat_param = self.fused.grad, self.fused.param       # peer-mapped buffers; m, v: this rank's shard
        else:
            self.flat_param = torch.zeros(n, dtype=torch.float32, device=device)
            self.flat_grad = torch.zeros_like(self.flat_param)
            self.opt = FlatAdam(self.flat_param, self.flat_grad, lr=lr)
        self.launches_per_update = 0 if not optimizer else (2 if self.fused is not None else 3)
        # the stand-in backward as a prepared C-ABI call (fixed gradient buffer; the loss outputs' pointers come per call)
        self._grad_fn, self._check, self._count = _ffi.lib().xa_grad_from_outputs_f32, _ffi.check, ops._count
        self._grad_tail = (ctypes.c_void_p(self.flat_grad.data_ptr()), self.flat_grad.numel())
        self._vp = ctypes.c_void_p
        self.update_events = None        # list of (start, end) CUDA events per update when --time-updates

    def forward(self, states, training=True):
        raise RuntimeError('the benchmark feeds complete rollouts: there is no rollout-time forward')

    def forward_into(self, states, actor_dst, critic_dst, i):
        if self.tables is not None:
            n = critic_dst.shape[0]
            actor_dst.copy_(self.tables[0][i, :n])
            critic_dst.copy_(self.tables[1][i, :n])

    def backward_and_step(self, d_actor, d_values, grad_norm=None):
        if not self.optimizer:
            return
        if self.update_events is not None:
            a, b = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
            a.record()
            self._update(d_actor, d_values, grad_norm)
            b.record()
            self.update_events.append((a, b))
            return
        self._update(d_actor, d_values, grad_norm)

    def _update(self, d_actor, d_values, grad_norm):
        comm, vp = self.comm, self._vp
        stream = vp(self.torch.cuda.current_stream(self.device).cuda_stream)
        n = d_values.shape[0]
        self._check('xa_grad_from_outputs_f32', self._grad_fn(vp(d_actor.data_ptr()), vp(d_values.data_ptr()), n, d_actor.shape[1],
                                                             *self._grad_tail, stream))
        self._count()
        self.step += 1
        if self.fused is not None:
            self.fused.step(self.step, grad_norm)
            self._count()
            return
        scale = 1.0
        if comm is not None and comm.world_size > 1 and self.c1 != 'none':
            comm.all_reduce_gradients_async(self.flat_grad)         # collective C1 (comm stream, high priority) ...
            comm.wait_gradients()                                   # ... and the optimiser waits for it
            scale = 1.0 / comm.world_size
        self.opt.step(grad_norm, scale)


def synthetic_host_rollout(T, E, shape, dtype, A, K, M, seed):
    """Scalars of one synthetic rollout on the host (SURVEY.md 8d distributions); frames are drawn on the device."""
    import numpy as np
    rng = np.random.default_rng(seed)
    N = T * E
    return {
        'rewards': rng.standard_normal((T, E)).astype(np.float32),
        'values': rng.standard_normal((T, E)).astype(np.float32),
        'last_values': rng.standard_normal(E).astype(np.float32),
        'dones': (rng.random((T + 1, E)) < 0.01).astype(np.float32),
        'actions': rng.integers(0, A, (T, E)).astype(np.float32),
        'log_probs': (-np.abs(rng.standard_normal((T, E))) - 0.5).astype(np.float32),
    }


def build_ppo(args, dev, comm, E, T, shape, dtype, A, n_params, rank, c1, network='stand-in'):
    """The drop-in PPO agent over fed rollouts, its buffers filled with one synthetic rollout resident in HBM."""
    import torch

    from xagents_b200 import feeds
    from xagents_b200.agents import PPO
    tdtype = torch.uint8 if dtype == 'uint8' else torch.float32
    envs = feeds.FedEnvs(E, shape, tdtype, A, device=dev)
    if network == 'nature-tc':
        from xagents_b200.agents import NatureCnnTc, TorchModel
        torch.manual_seed(0)
        net = TorchModel(NatureCnnTc(shape[-1], A).to(dev), comm=comm)
    else:
        net = SyntheticNetwork(dev, comm, n_params, optimizer=not args.no_optimizer, c1=c1)
    agent = PPO(envs, net, n_steps=T, quiet=True, device=dev)
    opts = dict(gather_mode=args.gather_mode, staging=args.staging, overlap=not args.no_overlap, sync=args.sync,
                dynamic=not args.static_gather, late_fork=args.late_fork)
    if args.gather_schedule:
        opts['gather_chunk'] = [int(x) for x in args.gather_schedule.split(',')]
    elif args.gather_chunk:
        opts['gather_chunk'] = args.gather_chunk
    if args.obs_gather != 'auto':
        opts['obs_gather'] = args.obs_gather == 'on'
    agent.pipeline_options = opts
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    if tdtype == torch.uint8:
        agent.ro_states.copy_(torch.randint(0, 256, agent.ro_states.shape, dtype=torch.uint8, device=dev, generator=gen))
    else:
        agent.ro_states.copy_(torch.randn(agent.ro_states.shape, device=dev, generator=gen))
    host = synthetic_host_rollout(T, E, shape, dtype, A, agent.ppo_epochs, agent.mini_batches, 1234 + rank)
    for k, attr in (('rewards', 'ro_rewards'), ('values', 'ro_values'), ('dones', 'ro_dones'), ('actions', 'ro_actions'),
                    ('log_probs', 'ro_log_probs')):
        getattr(agent, attr).copy_(torch.from_numpy(host[k]))
    last_values = torch.from_numpy(host['last_values']).to(dev)
    N = T * E
    perms = torch.stack([torch.randperm(N, device=dev, generator=gen).to(torch.int32) for _ in range(agent.ppo_epochs)])
    agent.rollout_source = lambda ag: last_values                 # the rollout is already in the agent's buffers
    hp = agent.hot_path()
    hp.perms.copy_(perms)
    agent.permutation_source = lambda epoch: hp.perms[epoch]      # identical permutation indices on every step, resident
    if network == 'stand-in':
        hp.actor_out.copy_(torch.randn(hp.actor_out.shape, device=dev, generator=gen))
        hp.critic_out.copy_(torch.randn(hp.critic_out.shape, device=dev, generator=gen))
    return agent, net, hp, host, perms, last_values


def timed_steps(step, stream, n_steps, comm, dev, sampler=None):
    """CUDA events on the launching stream, barrier + synchronize on both sides; max over ranks.  -> (ms total, [ms per step])"""
    import torch
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(n_steps + 1)]
    if comm is not None:
        comm.barrier()
    torch.cuda.synchronize(dev)
    if sampler:
        sampler.mark()
    marks[0].record(stream)
    t0 = time.perf_counter()
    for s in range(n_steps):
        step(s)
        marks[s + 1].record(stream)
    timed_steps.host_ms_per_step = (time.perf_counter() - t0) * 1e3 / n_steps      # time the HOST needed to issue a step
    torch.cuda.synchronize(dev)
    if comm is not None:
        comm.barrier()
    total = marks[0].elapsed_time(marks[-1])
    per_step = [marks[s].elapsed_time(marks[s + 1]) for s in range(n_steps)]
    if comm is not None:
        total = comm.max_over_ranks(total)
    return total, per_step


def h2d_roofline(dev, comm, pinned_obs):
    """Measured pinned-host -> device copy rate on this box: the denominator of the end-to-end number (it is PCIe-bound).
    `single`: rank 0 alone; `concurrent`: every rank at once (they share the host's memory system and PCIe root complexes)."""
    import torch
    flat = pinned_obs.view(-1)
    n = min(flat.numel(), 512 << 20)
    src = flat[:n]
    dst = torch.empty(n, dtype=torch.uint8, device=dev)

    def rate():
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize(dev)
        a.record()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        b.record()
        torch.cuda.synchronize(dev)
        return 3 * n / (a.elapsed_time(b) * 1e-3) / 1e9

    rank = comm.rank if comm is not None else 0
    single = rate() if rank == 0 else 0.0
    out = {'single_rank_GBs': single, 'bytes_per_copy': n}
    if comm is not None:
        comm.barrier()
        mine = rate()
        out['single_rank_GBs'] = comm.max_over_ranks(single)
        out['concurrent_min_rank_GBs'] = -comm.max_over_ranks(-mine)
        out['concurrent_sum_GBs'] = comm.sum_over_ranks([mine])[0]
    return out


def sharded_parity_check(comm, dev):
    """N>1, before timing: a 2-minibatch sharded train step through the prepared pipeline (C2 all-gather of moments included)
    against the single-process oracle on the equivalent global permutation (dist.global_minibatch_indices; reference semantics:
    xagents/ppo/agent.py:180-183 normalises with the statistics of the WHOLE minibatch).  Returns are checked bit-exactly, the
    staged minibatch rows bit-exactly, moments and the rank-averaged loss scalars to 1e-5.  The oracle is the checker only."""
    import numpy as np
    import torch

    import oracle
    from xagents_b200 import dist as xdist
    from xagents_b200 import hotpath, synthetic
    rank, world = comm.rank, comm.world_size
    T, A, K, MB = 16, 6, 2, 2
    E_total = 8 * world
    ro = synthetic.make_rollout(T, E_total, obs_shape=FRAME, n_actions=A, epochs=0, p_done=0.05)     # same arrays on every rank
    lo, hi = xdist.shard_range(E_total, rank, world)
    E = hi - lo
    n_local = E * T
    local_perms = [[np.random.default_rng(100 * g + k).permutation(n_local).astype(np.int32) for k in range(K)] for g in range(world)]
    glob = xdist.global_minibatch_indices(local_perms, T, E_total, world, MB)
    hp = hotpath.PPOHotPath(T, E, FRAME, A, ppo_epochs=K, mini_batches=MB, device=dev, comm=comm)
    for name in hp.ROLLOUT_FIELDS:
        getattr(hp, name).copy_(torch.from_numpy(np.ascontiguousarray(xdist.shard_time_major(getattr(ro, name), rank, world))))
    hp.perms.copy_(torch.from_numpy(np.stack(local_perms[rank])))
    b_local = n_local // MB
    for k in range(K):
        for m in range(MB):
            mine = glob[k][m][rank * b_local:(rank + 1) * b_local]          # this rank's part of global minibatch (k, m)
            hp.actor_out[k * MB + m].copy_(torch.from_numpy(ro.new_logits[mine]))
            hp.critic_out[k * MB + m].copy_(torch.from_numpy(ro.new_values[mine]))
    hp.prepare(torch.cuda.current_stream(dev))
    hp.run()
    torch.cuda.synchronize(dev)
    want = oracle.ppo_train_step(ro.obs, ro.rewards, ro.dones, ro.values, ro.last_values, ro.actions, ro.log_probs,
                                 [np.concatenate(glob[k]) for k in range(K)], ro.new_logits, ro.new_values, mini_batches=MB,
                                 keep_states=True)
    ok = np.array_equal(hp.returns.cpu().numpy(), want['returns'][:, lo:hi])
    flat_ret, flat_val = oracle.concat_step_batches(want['returns'], ro.values)
    all_moments = hp.all_moments.cpu().numpy()
    worst = 0.0
    sums = comm.sum_over_ranks(hp.scalars.cpu().numpy().astype(np.float64).reshape(-1).tolist())
    mean_scalars = np.asarray(sums).reshape(K * MB, 4) / world
    for i, mb in enumerate(want['minibatches']):
        adv = (flat_ret[mb['idx']] - flat_val[mb['idx']]).astype(np.float64)
        n, mean, std = xdist.combine_moments(all_moments[:, i])
        ok = ok and n == len(adv) and abs(mean - adv.mean()) <= 1e-9 and abs(std - adv.std()) <= 1e-9
        ref = np.array([mb['loss'], mb['pg'], mb['vl'], mb['entropy']], np.float64)
        scale = np.abs(ref).max()
        err = float(np.abs(mean_scalars[i] - ref).max() / scale)
        worst = max(worst, err)
        ok = ok and err <= 1e-5
        g, slot, row0 = hp._mb_place[i]
        if g >= hp.n_groups - hp.staging:                      # this minibatch's rows are still resident in a staging slot
            got = hp.mb_obs[slot, row0:row0 + b_local].cpu().numpy()
            ok = ok and np.array_equal(got, mb['states'][rank * b_local:(rank + 1) * b_local])
    all_ok = comm.max_over_ranks(0.0 if ok else 1.0) == 0.0
    return {'parity_checked': bool(all_ok), 'worst_scalar_rel_err': comm.max_over_ranks(worst),
            'what': f'{world}-rank sharded step (T={T}, n_envs={E_total}, {K} epochs x {MB} minibatches, 84x84x4 frames) vs the '
                    f'single-process oracle on the concatenated global permutation: returns and staged rows bit-exact, all-gathered '
                    f'moments and rank-averaged loss/pg/value/entropy <= 1e-5'}


def run_ppo(args):
    import numpy as np
    import torch

    from xagents_b200 import dist as xdist
    from xagents_b200 import feeds, ops

    rank, local_rank, world = xdist.init_from_env()
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device: xagents_b200 has no CPU path'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    comm = xdist.ShardComm(device=dev) if world > 1 else None
    name, total_envs, T, shape, dtype, A, desc = workload_of(args, world)
    lo, hi = xdist.shard_range(total_envs, rank, world)
    E = hi - lo
    n_params = CARTPOLE_MLP_PARAMS if name == 'c1' else NATURE_CNN_PARAMS
    c1 = args.c1
    if c1 == 'auto':
        c1 = 'nccl'
        if world > 1:
            from xagents_b200 import peer
            c1 = 'fused' if peer.available(comm) else 'nccl'
    parity = None
    if world > 1 and not args.no_parity_check:
        parity = sharded_parity_check(comm, dev)

    agent, net, hp, host, perms, last_values = build_ppo(args, dev, comm, E, T, shape, dtype, A, n_params, rank, c1, args.network)
    N, B, K, M = hp.N, hp.B, hp.K, hp.M
    stream = torch.cuda.current_stream(dev)
    n_gathers = hp.n_groups if hp.obs_gather else 0

    # the sampler starts before the warm-up (nvidia-smi, the fallback, takes ~100 ms to start); with NVML only the samples
    # taken inside the timed region are reported
    sampler = ClockSampler(local_rank) if (rank == 0 and not args.no_clock_sampler) else None
    if sampler:
        sampler.start()
    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        agent.train_step()
    torch.cuda.synchronize(dev)

    # ---- timed region ------------------------------------------------------------------------------
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_gathers)]
          for _ in range(args.steps)]
    cur = [0]

    def timed_gather(i, fn, fargs):
        a, b = ev[cur[0]][i]
        a.record(hp.data_stream)
        rc = fn(*fargs)
        b.record(hp.data_stream)
        return rc

    def step(s):
        cur[0] = s
        agent.train_step()

    hp.on_gather = timed_gather
    if args.time_updates and hasattr(net, 'update_events'):
        net.update_events = []
    ops.reset_launch_count()
    elapsed_ms, per_step = timed_steps(step, stream, args.steps, comm, dev, sampler)
    host_ms = timed_steps.host_ms_per_step
    hp.on_gather = None
    update_ms = None
    if getattr(net, 'update_events', None):
        update_ms = [a.elapsed_time(b) for a, b in net.update_events]
        net.update_events = None
    wrapper_launches = ops.launch_count()
    clocks = sampler.stop() if sampler else None
    gather_ms = [a.elapsed_time(b) for row in ev for (a, b) in row]
    ms_per_step = elapsed_ms / args.steps
    value = total_envs * T * args.steps / (elapsed_ms * 1e-3)
    assert agent.steps == (warmup + args.steps) * N, 'train_step() must advance agent.steps by n_steps * n_envs'
    agent.steps = 0
    launches = wrapper_launches                                  # every C-ABI launch goes through ops' counter (pipeline included)

    # ---- the metric's three stages alone (round-1 definition: nothing after the loss), for continuity ----------------
    bare = None
    if world == 1 and args.network == 'stand-in' and not args.no_optimizer:
        net.optimizer = False
        for _ in range(3):
            agent.train_step()
        ms_b, per_b = timed_steps(lambda s: agent.train_step(), stream, args.steps, None, dev)
        net.optimizer = True
        bare = {'value': total_envs * T * args.steps / (ms_b * 1e-3), 'ms_per_step': ms_b / args.steps,
                'ms_per_step_median': statistics.median(per_b),
                'what': 'the same train_step() with nothing after each loss (no gradient stand-in, no optimiser): GAE + gathers + losses only'}

    # ---- end to end: PPO.train_step() fed HOST rollouts ---------------------------------------------
    # feeds.HostRolloutFeed: every step's rollout, permutations and model outputs come from pinned host memory; rollout s+1
    # travels on a copy stream while train step s runs; the host reads the loss scalars of every step before going on.
    e2e = None
    if not args.no_e2e:
        pinned = {'obs': torch.empty(agent.ro_states.shape, dtype=agent.ro_states.dtype).pin_memory()}
        pinned['obs'].copy_(agent.ro_states)
        for k, v in host.items():
            pinned[k] = torch.from_numpy(v).pin_memory()
        pinned['perms'] = perms.cpu().pin_memory()
        extras = {'perms': (tuple(perms.shape), torch.int32)}
        if args.network == 'stand-in':
            pinned['actor_out'] = hp.actor_out.cpu().pin_memory()
            pinned['critic_out'] = hp.critic_out.cpu().pin_memory()
            extras.update(actor_out=(tuple(hp.actor_out.shape), torch.float32), critic_out=(tuple(hp.critic_out.shape), torch.float32))

        def on_switch(cur_set):
            if args.network == 'stand-in':
                net.tables = (cur_set['actor_out'], cur_set['critic_out'])

        feed = feeds.HostRolloutFeed(agent, lambda k: pinned, extras=extras, on_switch=on_switch)
        agent.rollout_source = feed
        agent.permutation_source = lambda epoch: feed.current['perms'][epoch]
        out_host = torch.empty(hp.scalars.shape, dtype=torch.float32).pin_memory()
        done = torch.cuda.Event()

        def e2e_run(n):
            last = 0.0
            for _ in range(n):
                agent.train_step()
                out_host.copy_(agent.hot_path().scalars, non_blocking=True)
                done.record(stream)
                done.synchronize()                           # the caller reads this step's losses
                last = float(out_host[0, 0])
            return last

        e2e_run(3)
        if comm is not None:
            comm.barrier()
        torch.cuda.synchronize(dev)
        n_e2e = max(4, min(args.steps, 20))
        t0 = time.perf_counter()
        e2e_run(n_e2e)
        torch.cuda.synchronize(dev)
        e_ms = (time.perf_counter() - t0) * 1e3              # host wall clock: includes every wait the caller sees
        if comm is not None:
            comm.barrier()
            e_ms = comm.max_over_ranks(e_ms)
        h2d = feed.h2d_bytes_per_rollout
        roof = h2d_roofline(dev, comm, pinned['obs'])
        link = roof.get('concurrent_min_rank_GBs', roof['single_rank_GBs'])
        e2e = {'value': total_envs * T * n_e2e / (e_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d,
               'd2h_bytes_per_step': out_host.numel() * 4, 'steps': n_e2e, 'ms_per_step': e_ms / n_e2e,
               'h2d_GBs_achieved_per_gpu': h2d / (e_ms / n_e2e * 1e-3) / 1e9, 'h2d_roofline': roof,
               'frac_of_h2d_roofline': (h2d / (e_ms / n_e2e * 1e-3) / 1e9) / link if link else None,
               'api': 'PPO.train_step() of the drop-in agent, rollout_source = feeds.HostRolloutFeed: rollout + permutations + model '
                      'outputs copied from pinned host buffers every step (uint8 frames; the next rollout travels on a copy stream '
                      'under the current step), loss scalars copied back and waited for every step'}
        del pinned, feed

    # ---- what the line needs from the pipeline, before it is torn down ---------------------------------
    alg = hp.algorithmic_bytes()
    info = {'group_sizes': list(hp.group_sizes), 'mean_group_rows': sum(hp.group_rows) / len(hp.group_rows), 'overlap': hp.overlap,
            'obs_mb': agent.ro_states.numel() * agent.ro_states.element_size() / 1e6, 'launches_per_update': getattr(net, 'launches_per_update', None),
            'c1': ('none (single GPU)' if world == 1 else c1), 'n_params': getattr(net, 'n_params', None)}
    graph_us = None
    if name == 'c1' and world == 1:
        # latency-bound shape: the bare pipeline (GAE, moments, gathers, losses) as ONE CUDA-graph replay
        replay = hp.capture()
        for _ in range(10):
            replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        a.record()
        for _ in range(200):
            replay()
        b.record()
        torch.cuda.synchronize(dev)
        graph_us = a.elapsed_time(b) / 200 * 1e3

    # ---- N>1: the same per-GPU work on ONE GPU without collectives, and C4 whole on one GPU (rank 0; the others wait) ----
    single = None
    if world > 1 and not args.no_single_gpu_compare and args.network == 'stand-in':
        del agent, net, hp
        torch.cuda.empty_cache()
        single = {}
        if rank == 0:
            for label, e_single in (('same_per_gpu_n_envs', E), ('whole_workload_on_one_gpu', total_envs)):
                ag1, net1, hp1, *_ = build_ppo(args, dev, None, e_single, T, shape, dtype, A, n_params, 0, 'nccl')
                for _ in range(3):
                    ag1.train_step()
                k = max(3, min(args.steps, 10))
                ms, _ = timed_steps(lambda s: ag1.train_step(), stream, k, None, dev)
                single[label] = {'n_envs': e_single, 'value': e_single * T * k / (ms * 1e-3), 'ms_per_step': ms / k, 'steps': k,
                                 'minibatches_per_gather_launch': list(hp1.group_sizes)}
                del ag1, net1, hp1
                torch.cuda.empty_cache()
        comm.barrier()

    if rank != 0:
        return
    peak, peak_src = hbm_peak()
    hp_obs_gather = bool(gather_ms)
    if not gather_ms:           # the network fetched the frames through the permutation: there is no gather launch to rate
        g_ms, achieved = None, 0.0
    else:
        g_ms = statistics.mean(gather_ms)
        achieved = alg['gather_per_launch'] / (g_ms * 1e-3) / 1e9
    traffic, traffic_source = None, None
    try:
        with open(os.path.join(ROOT, 'profiles', 'gather_traffic.json')) as f:
            traffic = json.load(f)['dram_bytes_per_row'] * info['mean_group_rows']
        traffic_source = ('profiles/gather_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of ONE ncu --set full capture of '
                          'gather_bulk_kernel at N=1 (32768 rows), per row moved, times the rows of this run\'s average launch -- '
                          'not measured in this run')
    except Exception:
        pass
    step_gbs = alg['total'] / (ms_per_step * 1e-3) / 1e9     # this rank's shard; ranks are symmetric
    med = statistics.median(per_step)
    meaningful = alg['total'] >= 32e6
    line = {
        'metric': METRIC if args.network == 'stand-in' else 'ppo_env_steps_per_sec_update_phase_with_network',
        'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': warmup,
        'ms_per_step': ms_per_step, 'ms_per_step_median': med, 'ms_per_step_min': min(per_step), 'ms_per_step_max': max(per_step),
        'host_issue_ms_per_step': host_ms, 'host_issue_note': 'wall time of the issuing loop / steps: an upper bound on interpreter time -- when a step has more launches than the launch queue lets the host run ahead, it includes waiting for the GPU', 'ms_per_step_each': [round(x, 4) for x in per_step], 'gather_launch_ms_each': [round(x, 3) for x in gather_ms],
        'value_at_median_step': N * 1e3 / med * world if world == 1 else None,
        'higher_is_better': True, 'scaling': 'weak' if world == 1 else 'strong',
        'vs_baseline': None, 'dtype': 'u8 rows + f32 scalars' if dtype == 'uint8' else 'f32', 'data': 'synthetic',
        'config': {'workload': desc, 'n_envs_per_gpu': E, 'samples_per_step_per_gpu': N, 'mini_batch_size_per_gpu': B,
                   'api': 'xagents_b200.agents.PPO.train_step() (the drop-in agent; run_ppo_epochs drives hotpath.PPOHotPath in place '
                          'over the agent\'s rollout buffers), rollout resident in HBM',
                   'gather_mode': args.gather_mode if hp_obs_gather else 'none: the first layer of the network reads every frame through the permutation (xa_nature_cnn_forward_indexed)',
                   'minibatches_per_gather_launch': info['group_sizes'] if hp_obs_gather else [],
                   'streams': 'gathers on a data stream, GAE/moments/losses/optimiser on the compute stream' if info['overlap'] else 'single stream',
                   'scalar_fields': 'read through the permutation inside the loss',
                   'after_each_loss': ('nothing (--no-optimizer)' if args.no_optimizer else
                                       (f'gradient stand-in kernel ({info["n_params"]} fp32 parameters) -> ' if args.network == 'stand-in'
                                        else 'Nature CNN backward on tcgen05 -> ') +
                                       ('' if world == 1 else {'nccl': 'NCCL all-reduce (comm stream), waited for by -> ',
                                                               'fused': 'ONE peer-memory kernel: reduce-scatter + global norm + clip + Adam on the '
                                                                        'rank\'s shard + all-gather of the updated weights; no NCCL; ',
                                                               'none': '(no all-reduce: diagnostic) -> '}[c1]) +
                                       ('' if (world > 1 and c1 == 'fused') else 'fused global-norm clip + Adam; ') +
                                       'the next minibatch\'s loss is ordered after it'),
                   'l2': f'inputs larger than L2: {info["obs_mb"]:.0f} MB of frames per GPU read once per epoch' if meaningful else
                         'rollout fits L2: latency-bound, not roofline-meaningful',
                   'collectives': ('none (single GPU)' if world == 1 else
                                   f'C1 = {c1}: gradients of {info["n_params"]} fp32 parameters per minibatch; C2 = 1 NCCL all-gather of advantage '
                                   f'moments per step')},
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': traffic, 'traffic_source': traffic_source, 'meaningful': meaningful and hp_obs_gather,
                     'kernel': 'gather_bulk_kernel' if (args.gather_mode != 'vector' and dtype == 'uint8') else 'gather_vector_kernel',
                     'peak_source': peak_src, 'algorithmic_bytes_per_launch': alg['gather_per_launch'],
                     'avg_launch_ms': g_ms, 'launches_timed': len(gather_ms),
                     'whole_step': {'algorithmic_bytes_per_step_per_gpu': alg['total'], 'achieved': step_gbs,
                                    'frac': step_gbs / peak, 'frac_of_nominal_8TBs': step_gbs / 8000.0,
                                    'achieved_at_median_step': alg['total'] / (med * 1e-3) / 1e9}},
        'gpu_launches': launches,
        'clocks': clocks,
    }
    if update_ms:
        line['update_ms'] = {'mean': statistics.mean(update_ms), 'median': statistics.median(update_ms), 'max': max(update_ms), 'n': len(update_ms),
                             'what': 'gradient stand-in + collective C1 + clip/Adam per minibatch, CUDA events on the compute stream inside the timed region'}
    if bare is not None:
        line['gae_gather_loss_only'] = bare
    if graph_us is not None:
        line['graph_replay_us_per_step'] = graph_us
        line['eager_us_per_step'] = ms_per_step * 1e3
    if parity is not None:
        line['parity_checked'] = parity['parity_checked']
        line['parity'] = parity
    if single is not None and single:
        line['single_gpu'] = single
        same = single['same_per_gpu_n_envs']['value']
        whole = single['whole_workload_on_one_gpu']['value']
        line['efficiency_same_per_gpu_E'] = value / (world * same)
        line['efficiency_vs_whole_workload_on_one_gpu'] = value / (world * whole)
    if e2e is not None:
        line['e2e'] = e2e
    if world == 1 and not args.no_cpu_baseline:
        _, base = cpu_reference(name, total_envs, T, shape, dtype, A, 3, 1)
        line['cpu_baseline'] = base
    emit(line)


def device_us(fn, reps, dev):
    """Average device time of back-to-back calls, in microseconds (CUDA events on the current stream)."""
    import torch
    for _ in range(5):
        fn()
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize(dev)
    return a.elapsed_time(b) / reps * 1e3


def graph_us(fn, reps, dev):
    """The same calls captured once into a CUDA graph of `reps` launches and replayed: device time without launch overhead."""
    import torch
    s = torch.cuda.Stream(dev)
    g = torch.cuda.CUDAGraph()
    torch.cuda.synchronize(dev)          # the inputs were produced on the caller's stream: the side stream must not race them
    with torch.cuda.stream(s):
        fn(s)
        s.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn(s)
    g.replay()
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record()
    torch.cuda.synchronize(dev)
    return a.elapsed_time(b) / (5 * reps) * 1e3


def gather_row_times(T, E, dev, reps):
    """xa_gather_rows (bulk / vector, prepared ctypes calls inside a CUDA graph: device time per launch) vs torch.index_select
    for the env-major reorder of T*E frames of 84x84x4 uint8."""
    import ctypes

    import torch

    from xagents_b200 import _ffi
    lib = _ffi.lib()
    N = T * E
    obs = torch.randint(0, 256, (T, E) + FRAME, dtype=torch.uint8, device=dev)
    dst = torch.empty((N,) + FRAME, dtype=torch.uint8, device=dev)
    perm = torch.randperm(N, device=dev).to(torch.int32)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    out = {}
    for mode, code in (('bulk', 1), ('vector', 2)):
        def call(s, code=code):
            if os.environ.get('XA_BENCH_TRACE'):
                print(f'gather_row_times T={T} E={E} mode={mode}', file=sys.stderr, flush=True)
            rc = lib.xa_gather_rows(P(obs), P(perm), P(dst), N, 28224, N, T, E, code, ctypes.c_void_p(s.cuda_stream))
            assert rc == 0, lib.xa_last_error()
        out[mode] = graph_us(call, reps, dev)
    flat, rows = obs.view(N, -1), ((perm.long() % T) * E + perm.long() // T)

    def sel(s):
        torch.index_select(flat, 0, rows, out=dst.view(N, -1))
    out['torch_index_select'] = graph_us(sel, reps, dev)
    return out


def run_a2c(args):
    """Config C2: A2C.train_step() of the drop-in agent on Pong-shaped rollouts (n_envs=16, n_steps=5): n-step returns scan,
    fused loss forward+backward over all T*E samples in time-major order (no flatten copy, no gather), gradient stand-in,
    fused clip+Adam.  80 samples: latency-bound, reported as microseconds per step."""
    import torch

    from xagents_b200 import feeds, ops
    from xagents_b200.agents import A2C
    assert int(os.environ.get('WORLD_SIZE', '1')) == 1, 'c2 is a single-GPU config'
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    name, E, T, shape, dtype, A, desc = workload_of(args, 1)
    envs = feeds.FedEnvs(E, shape, torch.uint8, A, device=dev)
    net = SyntheticNetwork(dev, None, NATURE_CNN_PARAMS, optimizer=not args.no_optimizer)
    agent = A2C(envs, net, n_steps=T, quiet=True, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234)
    agent.ro_states.copy_(torch.randint(0, 256, agent.ro_states.shape, dtype=torch.uint8, device=dev, generator=gen))
    host = synthetic_host_rollout(T, E, shape, dtype, A, 1, 1, 1234)
    for k, attr in (('rewards', 'ro_rewards'), ('values', 'ro_values'), ('dones', 'ro_dones'), ('actions', 'ro_actions')):
        getattr(agent, attr).copy_(torch.from_numpy(host[k]))
    last_values = torch.from_numpy(host['last_values']).to(dev)
    agent.rollout_source = lambda ag: last_values
    hp = agent.hot_path()
    hp.actor_out.copy_(torch.randn(hp.actor_out.shape, device=dev, generator=gen))
    hp.critic_out.copy_(torch.randn(hp.critic_out.shape, device=dev, generator=gen))
    stream = torch.cuda.current_stream(dev)
    sampler = ClockSampler(0)
    sampler.start()
    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        agent.train_step()
    steps = max(args.steps, 200)
    ops.reset_launch_count()
    elapsed_ms, per_step = timed_steps(lambda s: agent.train_step(), stream, steps, None, dev, sampler)
    launches = ops.launch_count()
    clocks = sampler.stop()
    N = T * E
    # the bare two-launch pipeline as a CUDA-graph replay (what a captured training loop would pay)
    replay_us = graph_us(lambda s: (hp.prepare(s), hp.run())[1], 50, dev)
    hp.prepare(stream)
    reorder = gather_row_times(T, E, dev, 50)
    # end to end: host rollout in, loss scalars out, every step
    pinned = {'obs': torch.empty(agent.ro_states.shape, dtype=torch.uint8).pin_memory()}
    pinned['obs'].copy_(agent.ro_states)
    for k in ('rewards', 'values', 'dones', 'actions', 'last_values'):
        pinned[k] = torch.from_numpy(host[k]).pin_memory()
    pinned['log_probs'] = torch.zeros((T, E)).pin_memory()
    feed = feeds.HostRolloutFeed(agent, lambda k: pinned)
    agent.rollout_source = feed
    out_host = torch.empty(4).pin_memory()
    done = torch.cuda.Event()

    def e2e_run(n):
        for _ in range(n):
            agent.train_step()
            out_host.copy_(agent.loss_scalars, non_blocking=True)
            done.record(stream)
            done.synchronize()
    e2e_run(5)
    torch.cuda.synchronize(dev)
    n_e2e = 100
    t0 = time.perf_counter()
    e2e_run(n_e2e)
    e_ms = (time.perf_counter() - t0) * 1e3
    alg_fused = 12 * N + 4 * E + (8 * A + 20) * N
    line = {
        'metric': 'a2c_env_steps_per_sec_returns_loss', 'value': N * steps / (elapsed_ms * 1e-3), 'unit': UNIT, 'n_gpus': 1, 'steps': steps,
        'warmup': warmup, 'ms_per_step': elapsed_ms / steps, 'ms_per_step_median': statistics.median(per_step), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': desc, 'api': 'xagents_b200.agents.A2C.train_step() (drop-in agent over hotpath.A2CHotPath), rollout resident in HBM',
                   'l2': 'rollout fits L2: latency-bound, not roofline-meaningful'},
        'eager_us_per_step': elapsed_ms / steps * 1e3, 'graph_replay_us_per_step_returns_plus_loss': replay_us,
        'env_major_reorder_80_frames_us': reorder,
        'roofline': {'bound': 'hbm', 'meaningful': False, 'achieved': alg_fused / (replay_us * 1e-6) / 1e9, 'peak': hbm_peak()[0], 'unit': 'GB/s',
                     'frac': alg_fused / (replay_us * 1e-6) / 1e9 / hbm_peak()[0], 'traffic': None,
                     'note': f'{alg_fused} algorithmic bytes per step (reorder fused away; 56 528 B/env-step if it were materialised): launch latency, not bandwidth'},
        'e2e': {'value': N * n_e2e / (e_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': feed.h2d_bytes_per_rollout, 'd2h_bytes_per_step': 16,
                'ms_per_step': e_ms / n_e2e, 'api': 'A2C.train_step() fed by feeds.HostRolloutFeed from pinned host buffers, loss scalars read back every step'},
        'gpu_launches': launches, 'clocks': clocks,
    }
    if not args.no_cpu_baseline:
        _, base = cpu_reference(name, E, T, shape, dtype, A, 20, 3)
        line['cpu_baseline'] = base
    emit(line)


def run_sweep(args):
    """Config C5: returns/GAE scan over n_steps 5..2048 x n_envs 16..65536 and the permute-gather over 80..524288 frames.
    Device time per launch of prepared ctypes calls (CUDA-graph replay for the small shapes); the CPU column is the oracle's
    NumPy loop (= the reference's own loop, xagents/ppo/agent.py:84-93)."""
    import ctypes

    import numpy as np
    import torch

    import oracle
    from xagents_b200 import _ffi
    assert int(os.environ.get('WORLD_SIZE', '1')) == 1, 'c5 is a single-GPU sweep'
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    lib = _ffi.lib()
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    peak, peak_src = hbm_peak()
    sampler = ClockSampler(0)
    sampler.start()
    sampler.mark()
    gae_rows, launches = [], 0
    for T in (5, 32, 128, 512, 2048):
        for E in (16, 256, 4096, 65536):
            rng = np.random.default_rng(T + E)
            r, v = rng.standard_normal((T, E)).astype(np.float32), rng.standard_normal((T, E)).astype(np.float32)
            lv, d = rng.standard_normal(E).astype(np.float32), (rng.random((T + 1, E)) < 0.01).astype(np.float32)
            rd, vd, lvd, dd = (torch.as_tensor(x).to(dev) for x in (r, v, lv, d))
            out = torch.empty((T, E), device=dev)
            nbytes = 16 * T * E + 4 * E
            reps = 50 if nbytes < 64e6 else 10

            def gae(s, mode=0):
                rc = lib.xa_gae_f32(P(rd), P(vd), P(lvd), P(dd), P(out), None, T, E, 0.99, 0.95, mode, ctypes.c_void_p(s.cuda_stream))
                assert rc == 0, lib.xa_last_error()

            def nstep(s):
                rc = lib.xa_nstep_returns_f32(P(rd), P(dd), P(lvd), P(out), T, E, 0.99, 0, ctypes.c_void_p(s.cuda_stream))
                assert rc == 0, lib.xa_last_error()
            t_auto = graph_us(gae, reps, dev)
            t_seq = graph_us(lambda s: gae(s, 1), reps, dev)
            t_n = graph_us(nstep, reps, dev)
            launches += 3 * 6 * reps
            t0 = time.perf_counter()
            n_cpu = 1 if T * E > 4e6 else 3
            for _ in range(n_cpu):
                oracle.gae_returns(r, d, v, lv, 0.99, 0.95)
            cpu_ms = (time.perf_counter() - t0) / n_cpu * 1e3
            gbs = nbytes / t_auto / 1e3
            gae_rows.append({'T': T, 'E': E, 'MB': nbytes / 1e6, 'auto_us': t_auto, 'sequential_bit_exact_us': t_seq, 'nstep_us': t_n,
                             'GBs': gbs, 'frac': gbs / peak, 'cpu_oracle_ms': cpu_ms, 'meaningful': nbytes >= 32e6})
    gather_rows = []
    for T, E in ((5, 16), (128, 16), (128, 64), (128, 256), (128, 1024), (128, 4096)):
        N = T * E
        nbytes = (2 * 28224 + 4) * N
        reps = 20 if nbytes < 1e9 else 4
        t = gather_row_times(T, E, dev, reps)
        launches += 2 * 6 * reps
        best = min(t['bulk'], t['vector'])
        gather_rows.append({'rows': N, 'MB': nbytes / 1e6, 'bulk_us': t['bulk'], 'vector_us': t['vector'],
                            'torch_index_select_us': t['torch_index_select'], 'bulk_GBs': nbytes / t['bulk'] / 1e3,
                            'frac': nbytes / t['bulk'] / 1e3 / peak, 'speedup_vs_torch': t['torch_index_select'] / best,
                            'meaningful': nbytes >= 100e6})
        torch.cuda.empty_cache()
    clocks = sampler.stop()
    top = gather_rows[-2]
    line = {'metric': 'gather_GBs_131072_frames', 'value': top['bulk_GBs'], 'unit': 'GB/s', 'n_gpus': 1, 'steps': 1, 'warmup': 0,
            'ms_per_step': top['bulk_us'] / 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8 rows + f32 scalars',
            'data': 'synthetic',
            'config': {'workload': 'c5: GAE / n-step scan sweep n_steps 5-2048 x n_envs 16-65536 (fp32) and permute-gather sweep 80-524288 '
                                   'frames of 84x84x4 uint8; device time per launch of prepared C-ABI calls (CUDA-graph replay)'},
            'roofline': {'bound': 'hbm', 'achieved': top['bulk_GBs'], 'peak': peak, 'unit': 'GB/s', 'frac': top['frac'], 'traffic': None,
                         'peak_source': peak_src, 'kernel': 'gather_bulk_kernel'},
            'gae_sweep': gae_rows, 'gather_sweep': gather_rows, 'gpu_launches': launches, 'clocks': clocks,
            'e2e': None, 'e2e_note': 'kernel sweep: no host-facing call'}
    if args.out:
        with open(args.out, 'w') as f:
            f.write('# C5 sweep: returns/GAE scan and permute-gather (python bench.py --workload c5)\n\n')
            f.write(f'Peak = {peak:.1f} GB/s ({peak_src}).  Device time per launch (CUDA-graph replay of prepared C-ABI calls).\n\n')
            f.write('| T | E | MB | auto us | GB/s | frac | sequential (bit-exact) us | n-step us | CPU oracle ms | speed-up | note |\n|---|---|---|---|---|---|---|---|---|---|---|\n')
            for r in gae_rows:
                f.write(f"| {r['T']} | {r['E']} | {r['MB']:.2f} | {r['auto_us']:.1f} | {r['GBs']:.0f} | {r['frac']:.2f} | {r['sequential_bit_exact_us']:.1f} | "
                        f"{r['nstep_us']:.1f} | {r['cpu_oracle_ms']:.2f} | {r['cpu_oracle_ms'] * 1e3 / r['auto_us']:.0f}x | "
                        f"{'' if r['meaningful'] else 'L2-resident, latency-bound'} |\n")
            f.write('\n| rows | MB moved | bulk us | bulk GB/s | frac | vector us | torch index_select us | best vs torch | note |\n|---|---|---|---|---|---|---|---|---|\n')
            for r in gather_rows:
                f.write(f"| {r['rows']} | {r['MB']:.1f} | {r['bulk_us']:.1f} | {r['bulk_GBs']:.0f} | {r['frac']:.2f} | {r['vector_us']:.1f} | "
                        f"{r['torch_index_select_us']:.1f} | {r['speedup_vs_torch']:.2f}x | {'' if r['meaningful'] else 'fits L2, latency-bound'} |\n")
    emit(line)


_RESULT_FD = None


def emit(line):
    """The ONE JSON line, on the process's original stdout."""
    data = (json.dumps(line) + '\n').encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    global _RESULT_FD
    args = parse_args()
    # stdout carries the ONE JSON line and nothing else: libraries that print there (NCCL announces its version on
    # stdout when a communicator is created) are pointed at stderr for the whole run
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == 'reference':
        run_reference(args)
        return
    if args.workload == 'c2':
        run_a2c(args)
    elif args.workload == 'c5':
        run_sweep(args)
    else:
        run_ppo(args)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
