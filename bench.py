#!/usr/bin/env python
"""bench.py -- headline benchmark of the nufhe_b200 engine (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--gate nand|mux] [--impl reference]

One "step" = one bootstrapped gate (default gate_nand) over a batch of B ciphertexts per GPU
(BASELINE.json configs[1]: B = 4096, n=500, N=1024, k=1, l=2, Bg=2^10, key switch t=8 base 4) on
synthetic data: seeded keys in the reference's RNG order, uniformly random LWE samples as operands
(the bootstrap does the same work whatever the plaintexts are).

  value      gates/s, whole job, operands resident in HBM when the timed region starts
  e2e        the same gate through the public API (vm.gate_nand) with HOST operands: pinned host ->
             device copies and the device -> host read of the result are inside the timed region
  roofline   the blind-rotate kernel against the measured HBM peak (MEASURED_PEAKS.json), bytes per
             SURVEY.md section 8(d)'s per-step model: n * (16384 * B + 65536) per launch
  cpu_baseline  the CPU oracle port (oracle/, C + OpenMP) on a bounded sample, all host cores

`--impl reference` times that CPU port alone (the reference is pure Python + JIT-compiled Reikna
kernels that cannot run here; its algorithm is restated in oracle/ and pinned to its own closures).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POLY = 1024
LWE_N = 500
SEED = 20260923


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=4096, help='ciphertexts per GPU per step')
    ap.add_argument('--gate', default='nand', choices=['nand', 'mux'])
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-sample', type=int, default=0, help='gates in the CPU sample (0 = auto)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


def host_cores():
    """Usable host threads: CPU affinity, capped by a cgroup CPU quota if the container has one."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        quota, period = open('/sys/fs/cgroup/cpu.max').read().split()
        if quota != 'max':
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except Exception:
        pass
    return n


def measured_peak_hbm():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def measured_traffic(batch):
    """dram__bytes_read + dram__bytes_write of one blind-rotate launch from the committed ncu capture
    (profiles/r1_traffic.json), only if it was taken at this batch size."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r1_traffic.json')) as f:
            d = json.load(f)
        if int(d['batch']) != int(batch):
            return None
        k = d['blind_rotate_kernel']
        return int(k['dram_bytes_read']) + int(k['dram_bytes_write'])
    except Exception:
        return None


def metric_name(args):
    return 'bootstrapped gates/sec (%s) at batch %d per GPU' % (args.gate.upper(), args.batch)


def workload_name(args):
    return ('gate_%s batch=%d/GPU, NTT transform, n=500 N=1024 k=1 l=2 Bg=2^10, keyswitch t=8 base=4, '
            'seeded keys (reference RNG order), uniform random LWE operands' % (args.gate, args.batch))


# --------------------------------------------------------------------------- CPU arm

def cpu_gate_sample(sample, gate):
    """Time the CPU oracle port on `sample` gates with all host threads.  Returns gates/s."""
    import numpy
    from oracle import oracle as O
    O.set_threads(host_cores())
    keys = cpu_gate_sample.keys
    if keys is None:
        keys = cpu_gate_sample.keys = O.OracleKeys(SEED)
    rng = numpy.random.RandomState(1)
    ops = [(rng.randint(-2**31, 2**31, size=(sample, LWE_N), dtype=numpy.int32),
            rng.randint(-2**31, 2**31, size=(sample,), dtype=numpy.int32)) for _ in range(3)]
    t = time.perf_counter()
    if gate == 'nand':
        O.gate_binary('nand', ops[0], ops[1], keys.bk, keys.ks)
    else:
        O.gate_mux(ops[0], ops[1], ops[2], keys.bk, keys.ks)
    dt = time.perf_counter() - t
    return sample / dt, dt


cpu_gate_sample.keys = None


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = host_cores()
    sample = args.cpu_sample or max(cores * 8, 16)
    for _ in range(args.warmup):
        cpu_gate_sample(min(sample, cores), args.gate)
    t_total = 0.0
    for _ in range(args.steps):
        _, dt = cpu_gate_sample(sample, args.gate)
        t_total += dt
    value = sample * args.steps / t_total
    line = {
        'impl': 'reference', 'metric': metric_name(args), 'value': value, 'unit': 'gates/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * t_total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'u64', 'data': 'synthetic',
        'config': {'workload': workload_name(args),
                   'note': 'CPU port of the reference algorithm (oracle/nufhe_oracle.c, OpenMP); each step '
                           'is a bounded sample of %d gates of the workload' % sample},
        'cpu_baseline': {'value': value, 'unit': 'gates/s', 'cores': cores, 'kind': 'port',
                         'sample': '%d gate_%s per step, %d steps' % (sample, args.gate, args.steps)},
        'e2e': {'value': value, 'unit': 'gates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- clocks sampler

class ClockSampler(threading.Thread):
    QUERY = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(
                    ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                     '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(',')]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = sorted(int(float(s[0])) for s in self.samples)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith('active') for s in self.samples)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': int(float(self.samples[0][1])), 'reasons': reasons,
                'samples': len(sm)}


# --------------------------------------------------------------------------- GPU arm

def run_b200_arm(args):
    import numpy
    import torch
    import torch.distributed as dist
    import nufhe_b200 as nufhe
    from nufhe_b200.lwe import LweSampleArray

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    ctx = nufhe.Context(rng=nufhe.DeterministicRNG(SEED), device_id=local_rank)
    thr = ctx.thread
    params = nufhe.NuFHEParameters()
    # Keys: rank 0 generates them (on its GPU, reference RNG order) and broadcasts the cloud key over
    # NCCL/NVLink once; there is no per-gate collective.
    if rank == 0:
        secret_key, cloud_key = ctx.make_key_pair()
    if world > 1:
        from nufhe_b200.api_low_level import NuFHECloudKey
        from nufhe_b200.bootstrap import BootstrapKey
        from nufhe_b200.tgsw import TransformedTGswSampleArray
        from nufhe_b200.lwe import LweKeyswitchKey
        if rank != 0:
            tg = TransformedTGswSampleArray.empty(thr, params.tgsw_params, (LWE_N,))
            ks_lwe = LweSampleArray.empty(thr, params.in_out_params, (N_POLY, 8, 4))
            cloud_key = NuFHECloudKey(params, BootstrapKey(params.in_out_params, tg), LweKeyswitchKey(ks_lwe))
        from nufhe_b200.sharding import cloud_key_tensors, broadcast_tensors
        broadcast_tensors(cloud_key_tensors(cloud_key), src=0)
        torch.cuda.synchronize()
    vm = ctx.make_virtual_machine(cloud_key)

    B = args.batch
    n_ops = 2 if args.gate == 'nand' else 3
    gen = torch.Generator(device='cpu').manual_seed(1234 + rank)
    host_ops = []
    for _ in range(n_ops):
        a = torch.randint(-2**31, 2**31, (B, LWE_N), generator=gen, dtype=torch.int64).to(torch.int32).pin_memory()
        b = torch.randint(-2**31, 2**31, (B,), generator=gen, dtype=torch.int64).to(torch.int32).pin_memory()
        host_ops.append((a, b))
    dev_ops = [LweSampleArray(params.in_out_params, a.to(thr.device), b.to(thr.device),
                              torch.zeros(B, dtype=torch.float32, device=thr.device)) for a, b in host_ops]
    dest = vm.empty_ciphertext((B,))
    gate = getattr(vm, 'gate_' + args.gate)
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=thr.device)   # > 126 MB L2
    launches_per_gate = 2          # fused bootstrap(s) + key switch, for NAND and for MUX alike

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        flush_buf.fill_(1)                      # L2 flush, inside the timed region (~0.1 ms)
        gate(*dev_ops, dest=dest)

    out_host_a = torch.empty((B, LWE_N), dtype=torch.int32).pin_memory()
    out_host_b = torch.empty((B,), dtype=torch.int32).pin_memory()

    def e2e_step():
        flush_buf.fill_(1)
        cts = [LweSampleArray(params.in_out_params, a.to(thr.device, non_blocking=True),
                              b.to(thr.device, non_blocking=True),
                              torch.zeros(B, dtype=torch.float32, device=thr.device)) for a, b in host_ops]
        r = gate(*cts)
        out_host_a.copy_(r.a, non_blocking=True)
        out_host_b.copy_(r.b, non_blocking=True)

    def timed(step_fn, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            step_fn()
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=thr.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    for _ in range(args.warmup):
        device_step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_total = timed(device_step, args.steps)
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)

    # the dominant kernel alone (blind rotate + extract), CUDA events on the launching stream
    from nufhe_b200.tgsw import engine_format
    bk_int = engine_format(thr, cloud_key.bootstrap_key.tgsw)
    ext = (thr.empty((B, N_POLY), torch.int32), thr.empty((B,), torch.int32))
    x1 = (dev_ops[0].a, dev_ops[0].b)
    x2 = (dev_ops[1].a, dev_ops[1].b)
    br_ms = []
    for i in range(args.warmup + args.steps):
        flush_buf.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        thr.bootstrap_extract(x1, x2, 2**29, -1, -1, 2**29, bk_int, out=ext)
        e1.record()
        torch.cuda.synchronize()
        if i >= args.warmup:
            br_ms.append(e0.elapsed_time(e1))
    br_avg_ms = sum(br_ms) / len(br_ms)
    ks_ms = []
    ks_arrays = cloud_key.keyswitch_key.device_arrays()
    for i in range(args.warmup + args.steps):
        flush_buf.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        thr.keyswitch(ks_arrays, ext, out=(dest.a, dest.b))
        e1.record()
        torch.cuda.synchronize()
        if i >= args.warmup:
            ks_ms.append(e0.elapsed_time(e1))
    ks_avg_ms = sum(ks_ms) / len(ks_ms)

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        alg_bytes = LWE_N * (16384 * B + 65536)
        achieved = alg_bytes / (br_avg_ms * 1e-3) / 1e9
        ms_per_step = ms_total / args.steps
        value = world * B * args.steps / (ms_total * 1e-3)
        e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
        h2d = world * n_ops * (B * LWE_N * 4 + B * 4)     # whole job, all ranks
        d2h = world * (B * LWE_N * 4 + B * 4)
        line = {
            'metric': metric_name(args), 'value': value, 'unit': 'gates/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'u64', 'data': 'synthetic',
            'config': {'workload': workload_name(args), 'global_batch': world * B,
                       'parallelism': 'ciphertext-sharded x%d, cloud key broadcast once over NCCL' % world,
                       'l2': 'flushed by a 256 MiB fill before every step (inside the timed region)',
                       'ms_per_gate': ms_per_step / B,
                       'published_reference_ms_per_gate': 0.35,
                       'published_note': 'nufhe README.md:64-65, NTT path, unnamed GPU and batch; not the same '
                                         'hardware/config, so vs_baseline stays null'},
            'e2e': {'value': e2e_value, 'unit': 'gates/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': launches_per_gate * args.steps,
            'clocks': sampler.summary(),
            'roofline': {'bound': 'hbm', 'kernel': 'blind_rotate_kernel', 'achieved': achieved, 'peak': peak,
                         'unit': 'GB/s', 'frac': achieved / peak, 'traffic': measured_traffic(B),
                         'algorithmic_bytes': alg_bytes,
                         'peak_source': peak_src, 'bytes_model': 'per-step: n*(16384*B+65536) per launch',
                         'ms_per_launch': br_avg_ms, 'share_of_step': br_avg_ms / ms_per_step,
                         'note': 'integer-issue bound, not HBM bound (SURVEY.md 8d); see profiles/'},
            'kernels_ms': {'blind_rotate_extract': br_avg_ms, 'keyswitch': ks_avg_ms},
            'build': thr.build_info(),
        }
        if not args.no_cpu_baseline:
            cores = host_cores()
            sample = args.cpu_sample or max(cores * 40, 64)
            cpu_gate_sample(min(sample, cores), args.gate)
            v, dt = cpu_gate_sample(sample, args.gate)
            line['cpu_baseline'] = {'value': v, 'unit': 'gates/s', 'cores': cores, 'kind': 'port',
                                    'sample': '%d gate_%s, %.1f s wall, oracle/nufhe_oracle.c + OpenMP'
                                              % (sample, args.gate, dt)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == '__main__':
    main()
