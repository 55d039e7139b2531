"""Import the UNMODIFIED reference (nucypher/nufhe, /root/reference) on a machine without
reikna / pycuda, to run its NumPy test closures (nufhe/*_cpu.py) as the tier-0 oracle.

Only used by tests/golden/make_golden.py *in the build container*; /root/reference does not
exist on the GPU box, so nothing under tests/ imports this module at test time.

Recipe: SURVEY.md Appendix D.  `reikna` is replaced by MagicMock modules (the *_cpu.py closures only
need `reikna.helpers.product` / `min_blocks`), and nufhe/transform/ntt_cpu.py:80 (`numpy.int32(val &
0xffffffff)`, which overflows on NumPy >= 2) is replaced by the equivalent wrapping conversion.
"""
import sys
import types
from unittest.mock import MagicMock

import numpy

REFERENCE_PATH = '/root/reference'


def _product(seq):
    r = 1
    for x in seq:
        r *= int(x)
    return r


def _min_blocks(length, block):
    return (length - 1) // block + 1


def import_reference():
    if 'nufhe' in sys.modules and getattr(sys.modules['nufhe'], '_is_reference_bridge', False):
        return sys.modules['nufhe']
    for name in ['reikna', 'reikna.cluda', 'reikna.cluda.api', 'reikna.cluda.dtypes',
                 'reikna.cluda.functions', 'reikna.core', 'reikna.algorithms',
                 'reikna.transformations']:
        sys.modules[name] = MagicMock()
    helpers = types.ModuleType('reikna.helpers')
    helpers.product = _product
    helpers.min_blocks = _min_blocks
    helpers.template_for = lambda *a, **k: MagicMock()
    helpers.log2 = lambda x: int(x).bit_length() - 1
    sys.modules['reikna.helpers'] = helpers
    sys.modules['reikna'].helpers = helpers
    for name in list(sys.modules):
        if name == 'nufhe' or name.startswith('nufhe.'):
            del sys.modules[name]
    sys.path.insert(0, REFERENCE_PATH)
    try:
        import nufhe
    finally:
        sys.path.remove(REFERENCE_PATH)
    nufhe._is_reference_bridge = True

    from nufhe.transform import ntt_cpu

    P = ntt_cpu.GaloisNumber.modulus

    def _gnum_to_i32(x):
        val = x.val
        lo = val & 0xffffffff
        if lo >= 2**31:
            lo -= 2**32
        r = lo - (1 if val > P // 2 else 0)
        if r < -2**31:
            r += 2**32
        return numpy.int32(r)

    ntt_cpu._gnum_to_i32 = _gnum_to_i32
    ntt_cpu.gnum_to_i32 = numpy.vectorize(_gnum_to_i32, otypes=[numpy.int32])
    return nufhe


if __name__ == '__main__':
    nufhe = import_reference()
    print('reference imported:', nufhe.__file__)
