"""Batch sweep of gate_nand (and gate_mux at 4096) plus the stand-alone transform, on one GPU.
Prints one JSON object; the committed copy lives in profiles/.  Usage: python tools/sweep.py [out.json]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                   # noqa: E402
import nufhe_b200 as nufhe                     # noqa: E402
from nufhe_b200.lwe import LweSampleArray      # noqa: E402

ctx = nufhe.Context(rng=nufhe.DeterministicRNG(20260923), device_id=0)
thr = ctx.thread
sk, ck = ctx.make_key_pair()
vm = ctx.make_virtual_machine(ck)
params = ck.params
gen = torch.Generator(device='cpu').manual_seed(99)
peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(
    os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=thr.device)


def rand_ct(B):
    a = torch.randint(-2**31, 2**31, (B, 500), generator=gen, dtype=torch.int64).to(torch.int32).to(thr.device)
    b = torch.randint(-2**31, 2**31, (B,), generator=gen, dtype=torch.int64).to(torch.int32).to(thr.device)
    return LweSampleArray(params.in_out_params, a, b, torch.zeros(B, dtype=torch.float32, device=thr.device))


def time_ms(fn, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


out = {'gate_nand': [], 'gate_mux': [], 'ntt': [], 'hbm_peak_gbs': peak, 'build': thr.build_info()}
BATCHES = [int(x) for x in os.environ.get('SWEEP_BATCHES', '1,64,256,592,768,1024,1536,2048,4096,16384,65536').split(',')]
for B in BATCHES:
    x, y = rand_ct(B), rand_ct(B)
    dest = vm.empty_ciphertext((B,))
    ms = time_ms(lambda: vm.gate_nand(x, y, dest=dest), 5 if B <= 4096 else 2)
    alg = 500 * (16384 * B + 65536)
    out['gate_nand'].append({'batch': B, 'ms': ms, 'ms_per_gate': ms / B, 'gates_per_s': B / ms * 1e3,
                             'hbm_gbs_per_step_model': alg / ms / 1e6, 'hbm_frac': alg / ms / 1e6 / peak})
    print(out['gate_nand'][-1], flush=True)
    del x, y, dest
for B in (() if os.environ.get('SWEEP_NO_MUX') else (4096,)):
    x, y, z = rand_ct(B), rand_ct(B), rand_ct(B)
    dest = vm.empty_ciphertext((B,))
    ms = time_ms(lambda: vm.gate_mux(x, y, z, dest=dest), 3)
    out['gate_mux'].append({'batch': B, 'ms': ms, 'ms_per_gate': ms / B, 'gates_per_s': B / ms * 1e3})
    print(out['gate_mux'][-1], flush=True)
    del x, y, z, dest
for NT in (() if os.environ.get('SWEEP_NO_MUX') else (4096, 65536, 262144)):
    polys = torch.randint(-2**31, 2**31, (NT, 1024), generator=gen, dtype=torch.int64).to(torch.int32).to(thr.device)
    f = thr.ntt_forward_i32(polys)
    outbuf = torch.empty_like(f)
    ib = torch.empty_like(polys)
    import ctypes
    def fwd():
        thr._call('nb_ntt_forward_i32', ctypes.c_void_p(polys.data_ptr()), ctypes.c_void_p(outbuf.data_ptr()), NT)
    def inv():
        thr._call('nb_ntt_inverse_i32', ctypes.c_void_p(f.data_ptr()), ctypes.c_void_p(ib.data_ptr()), NT)
    mf, mi = time_ms(fwd, 5), time_ms(inv, 5)
    out['ntt'].append({'transforms': NT, 'fwd_ms': mf, 'inv_ms': mi, 'fwd_gbs': NT * 12288 / mf / 1e6,
                       'inv_gbs': NT * 12288 / mi / 1e6, 'fwd_hbm_frac': NT * 12288 / mf / 1e6 / peak,
                       'inv_hbm_frac': NT * 12288 / mi / 1e6 / peak})
    print(out['ntt'][-1], flush=True)
    del polys, f, outbuf, ib
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], 'w'), indent=1)
