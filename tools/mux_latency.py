"""gate_mux and gate_nand latency at small batches (one GPU, L2 flushed before every call).
Usage: python tools/mux_latency.py [out.json]"""
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
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=thr.device)


def rand_ct(B):
    a = torch.randint(-2**31, 2**31, (B, 500), generator=gen, dtype=torch.int64).to(torch.int32).to(thr.device)
    b = torch.randint(-2**31, 2**31, (B,), generator=gen, dtype=torch.int64).to(torch.int32).to(thr.device)
    return LweSampleArray(params.in_out_params, a, b, torch.zeros(B, dtype=torch.float32, device=thr.device))


def time_ms(fn, reps=7):
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


out = []
for B in (1, 8, 16, 27, 32, 64):
    x, y, z = rand_ct(B), rand_ct(B), rand_ct(B)
    dest = vm.empty_ciphertext((B,))
    row = {'batch': B, 'gate_nand_ms': time_ms(lambda: vm.gate_nand(x, y, dest=dest)),
           'gate_mux_ms': time_ms(lambda: vm.gate_mux(x, y, z, dest=dest))}
    out.append(row)
    print(row, flush=True)
if len(sys.argv) > 1:
    json.dump({'rows': out, 'build': thr.build_info()}, open(sys.argv[1], 'w'), indent=1)
