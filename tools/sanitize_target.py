"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): a short blind rotation, a plain
external product, a key switch (both the one-wave and the split-j configuration) and the transforms."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                   # noqa: E402
from nufhe_b200.engine import Engine           # noqa: E402

eng = Engine(0)
g = torch.Generator(device='cpu').manual_seed(3)
r32 = lambda shape, lo=-2**31, hi=2**31: torch.randint(lo, hi, shape, generator=g, dtype=torch.int64).to(torch.int32).cuda()
rff = lambda shape: torch.randint(0, 2**62, shape, generator=g, dtype=torch.int64).cuda()
n = 3
bk = eng.bk_prepare(rff((n, 2, 2, 2, 1024)))
acc = r32((5, 2, 1024))
bara = r32((5, n), 0, 2048)
eng.blind_rotate(acc, bara, bk, return_accum=True)
eng.external_product(acc.clone(), bk, 1)
x1, x2 = (r32((3, n)), r32((3,))), (r32((3, n)), r32((3,)))
ext = eng.bootstrap_extract(x1, x2, 2**29, -1, -1, 2**29, bk)
ks = (r32((1024, 8, 4, 500)), r32((1024, 8, 4)), torch.zeros((1024, 8, 4)).cuda())
eng.keyswitch(ks, ext, want_cv=True)
big = (r32((300, 1024)), r32((300,)))
eng.keyswitch(ks, big, big, c=5)
p = r32((9, 1024))
eng.ntt_inverse_i32(eng.ntt_forward_i32(p))
# round 2: the work queue (more chains than resident CTAs, forced into 3 chunks with NUFHE_B200_FORCE_CHUNKS=3 so that
# accumulators are parked and resumed even with n = 3 steps), both CTA shapes; u64 transforms through the staging
# buffers (several sweeps per CTA); the LWE dot product and the key-switch-key kernel
B = int(os.environ.get('SANITIZE_BATCH', '700'))
xa, xb = (r32((B, n)), r32((B,))), (r32((B, n)), r32((B,)))
eng.bootstrap_extract(xa, xb, 2**29, -1, -1, 2**29, bk)
eng.bootstrap_extract2((xa, xb, 5, 1, 1), (xa, xb, 7, -1, 1), 2**29, bk)
pu = rff((2500, 1024))
eng.ntt_inverse_u64(eng.ntt_forward_u64(pu))
eng.ntt_inverse_i32(eng.ntt_forward_i32(r32((2500, 1024))))
key = r32((500,), 0, 2)
eng.lwe_dot(r32((77, 500)), key, add1=r32((77,)), add2=r32((77,)), sign=-1)
ks_a = torch.empty((16, 8, 4, 500), dtype=torch.int32).cuda()
ks_b = torch.empty((16, 8, 4), dtype=torch.int32).cuda()
ks_cv = torch.empty((16, 8, 4), dtype=torch.float32).cuda()
eng.make_keyswitch_key(ks_a, ks_b, ks_cv, r32((16,), 0, 2), key, r32((16, 8, 3, 500)), r32((16, 8, 3)), 2, 1e-9)
# the three CTA shapes of the fused kernel on the same small batch (the environment is read when an engine is created):
# throughput, wide (split inverse phases, exchange through shared memory) and wide2 (split forward phases as well)
# ... and the pair shape (a cluster of two CTAs per ciphertext, remote shared-memory stores)
for wide_max, wide2_max, pair_max in (('0', '0', '0'), ('1000000', '0', '0'), ('1000000', '1000000', '0'), ('0', '0', '1000000')):
    os.environ['NUFHE_B200_WIDE_MAX'], os.environ['NUFHE_B200_WIDE2_MAX'] = wide_max, wide2_max
    os.environ['NUFHE_B200_PAIR_MAX'] = pair_max
    e2 = Engine(0)
    e2.bootstrap_extract(x1, x2, 2**29, -1, -1, 2**29, bk)
    e2.blind_rotate(acc, bara, bk, return_accum=True)
    e2.external_product(acc.clone(), bk, 1)
    torch.cuda.synchronize()
torch.cuda.synchronize()
print('sanitize target done')
