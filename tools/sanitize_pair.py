"""compute-sanitizer target for the pair shape (blind_rotate_pair_kernel, both exchange variants): a gate bootstrap,
gate_mux's double job and a blind rotation with explicit accumulators, few LWE coefficients so that the tool finishes in
seconds.  Usage: compute-sanitizer --tool memcheck python tools/sanitize_pair.py"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                   # noqa: E402
from nufhe_b200.engine import Engine           # noqa: E402

N_LWE = int(os.environ.get('SANITIZE_N', '12'))
B = int(os.environ.get('SANITIZE_BATCH', '5'))
gen = torch.Generator(device='cpu').manual_seed(11)


def r32(shape, lo=-2**31, hi=2**31):
    return torch.randint(lo, hi, shape, generator=gen, dtype=torch.int64).to(torch.int32).cuda()


bk_ref = torch.randint(0, 2**62, (N_LWE, 2, 2, 2, 1024), generator=gen, dtype=torch.int64).cuda()
x1, x2 = (r32((B, N_LWE)), r32((B,))), (r32((B, N_LWE)), r32((B,)))
acc, bara = r32((B, 2, 1024)), r32((B, N_LWE), 0, 2048)
os.environ['NUFHE_B200_PAIR_MAX'] = '1000000'
for mode in ('1', '0'):
    os.environ['NUFHE_B200_PAIR_ASYNC'] = mode
    eng = Engine(0)
    bk = eng.bk_prepare(bk_ref)
    eng.bootstrap_extract(x1, x2, 2**29, -1, -1, 2**29, bk)
    eng.bootstrap_extract2((x1, x2, 5, 1, 1), (x1, x2, 7, -1, 1), 2**29, bk)
    eng.blind_rotate(acc, bara, bk, return_accum=True)
    torch.cuda.synchronize()
print('sanitize pair done')
