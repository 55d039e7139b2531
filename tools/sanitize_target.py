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
torch.cuda.synchronize()
print('sanitize target done')
