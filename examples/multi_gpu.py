"""Multi-GPU usage (counterpart of the reference's examples/multi_gpu.py, which uses one Python thread and
one Context per device and moves keys as pickled bytes).  Here: one process per GPU,

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 examples/multi_gpu.py

Rank 0 makes the keys and the ciphertexts; the cloud key is broadcast once (NCCL over NVLink), every
rank processes its contiguous shard of the ciphertext batch, and rank 0 gathers and decrypts.
"""
import os
import sys

import numpy
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nufhe_b200 as nufhe                                                  # noqa: E402
from nufhe_b200.api_low_level import NuFHECloudKey                           # noqa: E402
from nufhe_b200.bootstrap import BootstrapKey                                # noqa: E402
from nufhe_b200.lwe import LweKeyswitchKey, LweSampleArray                   # noqa: E402
from nufhe_b200.sharding import shard_bounds, cloud_key_tensors, broadcast_tensors, gather_shards   # noqa: E402
from nufhe_b200.tgsw import TransformedTGswSampleArray                       # noqa: E402

rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local_rank = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local_rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

ctx = nufhe.Context(rng=nufhe.DeterministicRNG(7), device_id=local_rank)
thr = ctx.thread
params = nufhe.NuFHEParameters()
size = 1024
rng = numpy.random.RandomState(1)
bits1, bits2 = rng.randint(0, 2, size).astype(bool), rng.randint(0, 2, size).astype(bool)

if rank == 0:
    secret_key, cloud_key = ctx.make_key_pair()
    ct1, ct2 = ctx.encrypt(secret_key, bits1), ctx.encrypt(secret_key, bits2)
else:
    tg = TransformedTGswSampleArray.empty(thr, params.tgsw_params, (500,))
    cloud_key = NuFHECloudKey(params, BootstrapKey(params.in_out_params, tg),
                              LweKeyswitchKey(LweSampleArray.empty(thr, params.in_out_params, (1024, 8, 4))))
    ct1, ct2 = (LweSampleArray.empty(thr, params.in_out_params, (size,)) for _ in range(2))
broadcast_tensors(cloud_key_tensors(cloud_key), src=0)          # ~98 MB, once
broadcast_tensors([ct1.a, ct1.b, ct2.a, ct2.b], src=0)          # the reference ships pickled slices instead

s, e = shard_bounds(size, world, rank)
vm = ctx.make_virtual_machine(cloud_key)
part = vm.gate_nand(ct1[s:e], ct2[s:e])
full_a = gather_shards(part.a, size, world, rank)
full_b = gather_shards(part.b, size, world, rank)
if rank == 0:
    result = LweSampleArray(params.in_out_params, full_a, full_b, torch.zeros(size, device=thr.device))
    assert (ctx.decrypt(secret_key, result) == ~(bits1 & bits2)).all()
    print('multi-GPU NAND over %d GPUs: %d bits ok' % (world, size))
dist.barrier()
dist.destroy_process_group()
