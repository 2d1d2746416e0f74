#!/usr/bin/env python
"""Headline benchmark: PPO env-steps/s through GAE + permute-gather + loss (BASELINE.json metric).

    python bench.py [--gpus N --steps K --warmup W]            our arm (one process per GPU under torchrun)
    python bench.py --impl reference [--gpus N ...]            the reference's CPU path on the host cores

A "step" is one PPO train step of the hot path over one synthetic rollout: 1 GAE scan, the advantage
moments of all K*M minibatches, then K*M x (minibatch gather of uint8 frames + 4 scalar fields, fused
loss forward+backward).  N=1 runs config C3 (n_envs=256, n_steps=128, 4 epochs x 4 minibatches, 84x84x4
uint8 frames); N>1 runs C4 (n_envs=4096 sharded over the ranks, NCCL gradient all-reduce per minibatch
on a side stream + one all-gather of advantage moments per step).  Prints ONE JSON line on rank 0.
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
FALLBACK_HBM_GBS = 6650.0              # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default=None, choices=[None, 'c3', 'c4'])
    ap.add_argument('--n-envs', type=int, default=None, help='total n_envs (overrides the workload default)')
    ap.add_argument('--n-steps', type=int, default=128)
    ap.add_argument('--gather-mode', default='auto', choices=['auto', 'bulk', 'vector'])
    ap.add_argument('--scan-mode', default='auto', choices=['auto', 'sequential', 'chunked'])
    ap.add_argument('--no-overlap', action='store_true', help='gathers on the compute stream instead of a data stream')
    ap.add_argument('--materialize-fields', action='store_true', help='gather the 4 scalar fields instead of reading them through idx')
    ap.add_argument('--staging', type=int, default=2)
    ap.add_argument('--gather-chunk', type=int, default=None, help='minibatches per gather launch (default: one epoch)')
    ap.add_argument('--gather-schedule', default=None, help='explicit launch schedule, e.g. 4,4,4,3,1')
    ap.add_argument('--no-grad-allreduce', action='store_true', help='diagnostic: drop collective C1 (gradient all-reduce per minibatch)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-sample-envs', type=int, default=64)
    return ap.parse_args()


def hbm_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return FALLBACK_HBM_GBS, 'fallback (B200_PROFILING.md)'


def workload_of(args, world):
    name = args.workload or ('c3' if world == 1 else 'c4')
    total_envs = args.n_envs or (256 if name == 'c3' else 4096)
    desc = (f'{name}: PPO on synthetic Atari frames (84x84x4 uint8), n_envs={total_envs}'
            f'{" sharded over %d ranks" % world if world > 1 else ""}, n_steps={args.n_steps}, '
            f'4 epochs x 4 minibatches, 6 actions')
    return name, total_envs, desc


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
                power = nv.nvmlDeviceGetPowerUsage(handle) / 1000.0
                bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                self.samples.append((float(sm), power, [n for n, m in masks if bits & m]))
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
            return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': self.sm_max,
                    'power_w_max': max((x[1] for x in used), default=None), 'samples': len(sm), 'reasons': reasons,
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
def cpu_reference(args, sample_envs, steps, warmup):
    """The reference's CPU path (oracle port; TF is not installable here) on a bounded sample."""
    import torch

    # all the host threads the box has: torchrun exports OMP_NUM_THREADS=1, which would make the N>1 reference arm 4x slower
    # than the N=1 one for no reason of the reference's
    try:
        n_threads = len(os.sched_getaffinity(0))
    except AttributeError:
        n_threads = os.cpu_count() or 1
    torch.set_num_threads(max(1, n_threads))

    from oracle import cpu_path
    from xagents_b200 import synthetic
    ro = synthetic.make_rollout(args.n_steps, sample_envs, epochs=4)
    res = cpu_path.time_cpu_baseline(ro, steps=steps, warmup=warmup, layout='reference')
    sample = (f'{steps} timed train steps (after {warmup} warm-up) of n_envs={sample_envs} x n_steps={args.n_steps} '
              f'(= {sample_envs * args.n_steps} samples/step, fp32 observations as the reference stores them), '
              f'4 epochs x 4 minibatches; env-steps/s is size-independent on the CPU')
    # also with the observations kept uint8 on the host (a quarter of the bytes the reference's fp32 layout
    # moves) -- the fairest CPU number for the byte movement itself; `value` stays the reference's own layout
    res_u8 = cpu_path.time_cpu_baseline(ro, steps=max(1, steps // 2), warmup=1, layout='uint8')
    return res, {'value': res['env_steps_per_sec'], 'unit': UNIT, 'cores': res['threads'], 'kind': 'port', 'sample': sample,
                 'value_uint8_obs': res_u8['env_steps_per_sec'], 'host_cpus': os.cpu_count(),
                 'torch_threads': torch.get_num_threads()}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if rank != 0:
        return
    name, total_envs, desc = workload_of(args, max(world, args.gpus))
    # bound the whole run to a few minutes: ~40 ms of CPU work per env of sample per step
    budget_s = 150.0
    per_env_s = 0.04
    sample = int(min(args.cpu_sample_envs, max(8, budget_s / ((args.steps + args.warmup) * per_env_s))))
    sample -= sample % 4
    res, base = cpu_reference(args, max(sample, 4), args.steps, args.warmup)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': res['env_steps_per_sec'], 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': res['seconds_per_step'] * 1e3,
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
def run_ours(args):
    import numpy as np
    import torch

    from xagents_b200 import dist as xdist
    from xagents_b200 import hotpath

    rank, local_rank, world = xdist.init_from_env()
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device: xagents_b200 has no CPU path'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    comm = xdist.ShardComm(device=dev) if world > 1 else None
    name, total_envs, desc = workload_of(args, world)
    lo, hi = xdist.shard_range(total_envs, rank, world)
    E, T, A = hi - lo, args.n_steps, 6
    hp = hotpath.PPOHotPath(T, E, (84, 84, 4), A, device=dev, gather_mode=args.gather_mode, scan_mode=args.scan_mode,
                            comm=comm, fuse_fields=not args.materialize_fields, staging=args.staging,
                            overlap=not args.no_overlap,
                            gather_chunk=[int(x) for x in args.gather_schedule.split(',')] if args.gather_schedule else args.gather_chunk)
    N, B, K, M = hp.N, hp.B, hp.K, hp.M

    # ---- synthetic rollout, resident in HBM before the timed region ---------------------------------
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    hp.obs.copy_(torch.randint(0, 256, hp.obs.shape, dtype=torch.uint8, device=dev, generator=gen))
    rng = np.random.default_rng(1234 + rank)
    host = {
        'rewards': rng.standard_normal((T, E)).astype(np.float32),
        'values': rng.standard_normal((T, E)).astype(np.float32),
        'last_values': rng.standard_normal(E).astype(np.float32),
        'dones': (rng.random((T + 1, E)) < 0.01).astype(np.float32),
        'actions': rng.integers(0, A, (T, E)).astype(np.float32),
        'log_probs': (-np.abs(rng.standard_normal((T, E))) - 0.5).astype(np.float32),
    }
    for k, v in host.items():
        getattr(hp, k).copy_(torch.from_numpy(v))
    for k in range(K):
        hp.perms[k].copy_(torch.randperm(N, device=dev, generator=gen).to(torch.int32))
    hp.actor_out.copy_(torch.randn(hp.actor_out.shape, device=dev, generator=gen))
    hp.critic_out.copy_(torch.randn(hp.critic_out.shape, device=dev, generator=gen))
    grad_buf = torch.zeros(NATURE_CNN_PARAMS, device=dev) if comm is not None else None
    stream = torch.cuda.current_stream(dev)
    hp.prepare(stream)
    n_gathers = hp.n_groups
    with_c1 = comm is not None and not args.no_grad_allreduce
    after_loss = (lambda i: comm.all_reduce_gradients_async(grad_buf)) if with_c1 else None

    def step(on_gather=None):
        hp.run(on_gather=on_gather, after_loss=after_loss)
        if comm is not None:
            comm.wait_gradients()

    # the sampler starts before the warm-up (nvidia-smi, the fallback, takes ~100 ms to start); with NVML only the samples
    # taken inside the timed region are reported
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize(dev)

    # ---- timed region: CUDA events on the launching stream, barrier + synchronize on both sides ------
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_gathers)]
          for _ in range(args.steps)]
    cur = [0]

    def timed_gather(i, fn, fargs):
        a, b = ev[cur[0]][i]
        a.record(hp.data_stream)
        rc = fn(*fargs)
        b.record(hp.data_stream)
        return rc

    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if comm is not None:
        comm.barrier()
    torch.cuda.synchronize(dev)
    if sampler:
        sampler.mark()
    start.record(stream)
    for s in range(args.steps):
        cur[0] = s
        step(timed_gather)
    stop.record(stream)
    torch.cuda.synchronize(dev)
    if comm is not None:
        comm.barrier()
    clocks = sampler.stop() if sampler else None
    elapsed_ms = start.elapsed_time(stop)
    if comm is not None:
        elapsed_ms = comm.max_over_ranks(elapsed_ms)
    gather_ms = [a.elapsed_time(b) for row in ev for (a, b) in row]
    ms_per_step = elapsed_ms / args.steps
    value = total_envs * T * args.steps / (elapsed_ms * 1e-3)

    # ---- end to end: host buffers in, loss scalars out, every step -----------------------------------
    # Two device-side rollout slots: the copy stream uploads step s+1 from pinned host memory while the compute
    # streams work on step s; the host reads the loss scalars of every step (D2H + event wait) before going on.
    e2e = None
    if not args.no_e2e:
        pinned = {'obs': torch.empty(hp.obs.shape, dtype=torch.uint8).pin_memory()}
        pinned['obs'].copy_(hp.obs)
        for k, v in host.items():
            pinned[k] = torch.from_numpy(v).pin_memory()
        pinned['perms'] = hp.perms.cpu().pin_memory()
        pinned['actor_out'] = hp.actor_out.cpu().pin_memory()
        pinned['critic_out'] = hp.critic_out.cpu().pin_memory()
        h2d = sum(t.numel() * t.element_size() for t in pinned.values())
        hp2 = hotpath.PPOHotPath(T, E, (84, 84, 4), A, device=dev, gather_mode=args.gather_mode, scan_mode=args.scan_mode,
                                 comm=comm, fuse_fields=hp.fuse_fields, staging=args.staging, overlap=hp.overlap,
                                 gather_chunk=hp.group_sizes).prepare(stream)
        slots = [hp, hp2]
        out_host = [torch.empty(hp.scalars.shape, dtype=torch.float32).pin_memory() for _ in slots]
        d2h = out_host[0].numel() * 4
        copy_stream = torch.cuda.Stream(dev)
        uploaded = [torch.cuda.Event() for _ in slots]
        finished = [torch.cuda.Event() for _ in slots]

        def upload(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(finished[i])          # slot i's previous step no longer reads its buffers
                for k, t in pinned.items():
                    getattr(slots[i], k).copy_(t, non_blocking=True)
                uploaded[i].record(copy_stream)

        def e2e_run(n_steps_e2e):
            for ev_ in finished:
                ev_.record(stream)
            upload(0)
            last = 0.0
            for s_ in range(n_steps_e2e):
                i = s_ & 1
                if s_ + 1 < n_steps_e2e:
                    upload(i ^ 1)                            # next step's inputs travel under this step's kernels
                stream.wait_event(uploaded[i])
                slots[i].run(after_loss=after_loss)
                if comm is not None:
                    comm.wait_gradients()
                out_host[i].copy_(slots[i].scalars, non_blocking=True)
                finished[i].record(stream)
                finished[i].synchronize()                    # the caller reads this step's losses
                last = float(out_host[i][0, 0])
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
        e2e = {'value': total_envs * T * n_e2e / (e_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d,
               'd2h_bytes_per_step': d2h, 'steps': n_e2e, 'ms_per_step': e_ms / n_e2e,
               'api': 'PPOHotPath: rollout + permutations + model outputs copied from pinned host buffers (double-buffered, '
                      'copy stream), run(), loss scalars read back and waited for every step'}
        del pinned, hp2, slots

    if rank != 0:
        return
    peak, peak_src = hbm_peak()
    alg = hp.algorithmic_bytes()
    g_ms = statistics.mean(gather_ms)
    achieved = alg['gather_per_launch'] / (g_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'gather_traffic.json')) as f:
            traffic = json.load(f)['dram_bytes_per_row'] * (sum(hp.group_rows) / len(hp.group_rows))   # ncu, per row moved
    except Exception:
        pass
    step_gbs = alg['total'] / (ms_per_step * 1e-3) / 1e9     # this rank's shard; ranks are symmetric
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak' if world == 1 else 'strong',
        'vs_baseline': None, 'dtype': 'u8 rows + f32 scalars', 'data': 'synthetic',
        'config': {'workload': desc, 'n_envs_per_gpu': E, 'samples_per_step_per_gpu': N, 'mini_batch_size_per_gpu': B,
                   'gather_mode': args.gather_mode, 'scan_mode': args.scan_mode,
                   'minibatches_per_gather_launch': hp.group_sizes,
                   'streams': 'gathers on a data stream, GAE/moments/losses on the compute stream' if hp.overlap else 'single stream',
                   'scalar_fields': 'read through the permutation inside the loss' if hp.fuse_fields else 'gathered per minibatch',
                   'l2': f'inputs larger than L2: {hp.obs.numel() / 1e6:.0f} MB of frames per GPU read once per epoch',
                   'collectives': ('none (single GPU)' if world == 1 else
                                   f'NCCL all-reduce of {NATURE_CNN_PARAMS} fp32 gradients per minibatch on a side stream '
                                   f'+ 1 all-gather of advantage moments per step')},
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': traffic, 'kernel': 'gather_bulk_kernel' if args.gather_mode != 'vector' else 'gather_vector_kernel',
                     'peak_source': peak_src, 'algorithmic_bytes_per_launch': alg['gather_per_launch'],
                     'avg_launch_ms': g_ms, 'launches_timed': len(gather_ms),
                     'whole_step': {'algorithmic_bytes_per_step_per_gpu': alg['total'], 'achieved': step_gbs,
                                    'frac': step_gbs / peak, 'frac_of_nominal_8TBs': step_gbs / 8000.0}},
        'gpu_launches': hp.kernel_launches_per_step * args.steps,
        'clocks': clocks,
    }
    if e2e is not None:
        line['e2e'] = e2e
    if world == 1 and not args.no_cpu_baseline:
        _, base = cpu_reference(args, args.cpu_sample_envs, 3, 1)
        line['cpu_baseline'] = base
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
    else:
        run_ours(args)
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
