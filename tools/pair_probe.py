"""Quick probe of the pair shape (one ciphertext on a cluster of two SMs): gate_nand on a small batch in every shape,
outputs compared with the throughput shape.  Usage: python tools/pair_probe.py [batch]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                   # noqa: E402
from nufhe_b200.engine import Engine           # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
gen = torch.Generator(device='cpu').manual_seed(5)


def r32(shape, lo=-2**31, hi=2**31):
    return torch.randint(lo, hi, shape, generator=gen, dtype=torch.int64).to(torch.int32).cuda()


outs = {}
bk_ref = torch.randint(0, 2**62, (500, 2, 2, 2, 1024), generator=gen, dtype=torch.int64).cuda()
x1, x2 = (r32((B, 500)), r32((B,))), (r32((B, 500)), r32((B,)))
for name, env in (('default', ('0', '0', '0', '1')), ('pair_barrier', ('0', '0', '1000000', '0')), ('pair_async', ('0', '0', '1000000', '1'))):
    for k, v in zip(('NUFHE_B200_WIDE_MAX', 'NUFHE_B200_WIDE2_MAX', 'NUFHE_B200_PAIR_MAX', 'NUFHE_B200_PAIR_ASYNC'), env):
        os.environ[k] = v
    eng = Engine(0)
    bk = eng.bk_prepare(bk_ref)
    a, b = eng.bootstrap_extract(x1, x2, 2**29, -1, -1, 2**29, bk)
    torch.cuda.synchronize()
    outs[name] = (a.cpu(), b.cpu())
    ok = all((x == y).all().item() for x, y in zip(outs['default'], outs[name]))
    print(name, 'equal to default:', ok, flush=True)
    assert ok, name
print('pair probe ok')
