"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun, NCCL).  Rank 0 generates the seeded keys on its
GPU and broadcasts the cloud key (nufhe_b200.sharding, the bench's set-up path); EVERY rank then runs gate_nand and
gate_mux on its own seeded operands with the key it received and compares all outputs with the CPU oracle (whose
seeded keys are the same keys).  Mirrors the reference's examples/multi_gpu.py:86-104."""
import os
import sys

import numpy
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import gen_inputs as G                                                        # noqa: E402
from oracle import oracle as O                                                # noqa: E402
import nufhe_b200 as nufhe                                                    # noqa: E402
from nufhe_b200.api_low_level import NuFHECloudKey                            # noqa: E402
from nufhe_b200.bootstrap import BootstrapKey                                 # noqa: E402
from nufhe_b200.lwe import LweKeyswitchKey, LweSampleArray                    # noqa: E402
from nufhe_b200.sharding import cloud_key_tensors, broadcast_tensors, shard_bounds, gather_shards   # noqa: E402
from nufhe_b200.tgsw import TransformedTGswSampleArray                        # noqa: E402

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
local_rank = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local_rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
ctx = nufhe.Context(rng=nufhe.DeterministicRNG(G.GATE_SEED), device_id=local_rank)
thr = ctx.thread
params = nufhe.NuFHEParameters()
if rank == 0:
    secret_key, cloud_key = ctx.make_key_pair()
else:
    tg = TransformedTGswSampleArray.empty(thr, params.tgsw_params, (500,))
    tg.samples.a.coeffs.zero_()
    ks = LweSampleArray.empty(thr, params.in_out_params, (1024, 8, 4))
    ks.a.zero_(); ks.b.zero_(); ks.current_variances.zero_()
    cloud_key = NuFHECloudKey(params, BootstrapKey(params.in_out_params, tg), LweKeyswitchKey(ks))
broadcast_tensors(cloud_key_tensors(cloud_key), src=0)
torch.cuda.synchronize()
vm = ctx.make_virtual_machine(cloud_key)

keys = O.OracleKeys(G.GATE_SEED)                     # same seed, same RNG order: the same keys on the CPU
B = 24
rng = numpy.random.RandomState(1000 + rank)          # every rank works on DIFFERENT ciphertexts
bits = [rng.randint(0, 2, B).astype(bool) for _ in range(3)]
cts = [keys.encrypt(b) for b in bits]


def dev(ct):
    return LweSampleArray(params.in_out_params, thr.to_device(ct[0]), thr.to_device(ct[1]),
                          torch.zeros(B, dtype=torch.float32, device=thr.device))


d = [dev(c) for c in cts]
r = vm.gate_nand(d[0], d[1])
want = O.gate_binary('nand', cts[0], cts[1], keys.bk, keys.ks)
ok = bool((r.a.cpu().numpy() == want[0]).all() and (r.b.cpu().numpy() == want[1]).all())
m = vm.gate_mux(d[0], d[1], d[2])
want = O.gate_mux(cts[0], cts[1], cts[2], keys.bk, keys.ks)
ok = ok and bool((m.a.cpu().numpy() == want[0]).all() and (m.b.cpu().numpy() == want[1]).all())
ok = ok and bool((keys.decrypt((m.a.cpu().numpy(), m.b.cpu().numpy())) == numpy.where(bits[0], bits[1], bits[2])).all())
# the gather path of nufhe_b200.sharding over NCCL: every rank's slice lands where shard_bounds says
s, e = shard_bounds(B * world, world, rank)
ga = gather_shards(r.a, B * world, world, rank)
assert ga.shape[0] == B * world and torch.equal(ga[s:e], r.a)
flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=thr.device)
gathered = [torch.zeros_like(flag) for _ in range(world)]
dist.all_gather(gathered, flag)
print('rank %d parity %s' % (rank, 'ok' if ok else 'MISMATCH'), flush=True)
if rank == 0:
    print('ranks ok: %s' % [int(g.item()) for g in gathered], flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
