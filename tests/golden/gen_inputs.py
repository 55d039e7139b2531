"""Seeded input generators shared by tests/golden/make_golden.py (which feeds them to the REFERENCE's
NumPy closures and stores the outputs) and by the tests (which feed the same inputs to the oracle and
to the CUDA path).  numpy.random.RandomState streams are stable across NumPy versions, so only seeds
and expected outputs need to be committed -- not the 65 MB key-switch key.

Ranges mirror the reference's tests (test/utils.py:41-57, test/test_tgsw.py, test/test_lwe.py, ...).
"""
import numpy

P = 2**64 - 2**32 + 1
N = 1024


def rs(seed):
    return numpy.random.RandomState(seed)


def ff_numbers(rng, shape):
    """uniform in [0, p) ('ff_number', test/utils.py:31-33)"""
    hi = rng.randint(0, 2**32, size=shape, dtype=numpy.uint64)
    lo = rng.randint(0, 2**32, size=shape, dtype=numpy.uint64)
    v = (hi << numpy.uint64(32)) | lo
    return numpy.where(v >= numpy.uint64(P), v - numpy.uint64(P), v).astype(numpy.uint64)


def torus32(rng, shape, lo=-2**31, hi=2**31):
    return rng.randint(lo, hi, size=shape, dtype=numpy.int32)


# Field-arithmetic known-answer operands: the regression values of test/test_transform/test_arithmetic.py
FF_EDGE = numpy.array([
    0, 1, 2, 2**32 - 1, 2**32, 2**32 + 1, 2**63, P - 1, P - 2, P // 2, P // 2 + 1,
    11509900421665959066, (P - 1) - ((P - 1) % 2**33), 0xfffffffe00000001, 0xffffffff], numpy.uint64)


def arithmetic_inputs():
    rng = rs(101)
    a = numpy.concatenate([numpy.repeat(FF_EDGE, FF_EDGE.size), ff_numbers(rng, (512,))])
    b = numpy.concatenate([numpy.tile(FF_EDGE, FF_EDGE.size), ff_numbers(rng, (512,))])
    s = rng.randint(0, 192, size=a.shape).astype(numpy.uint32)
    s[:192] = numpy.arange(192)
    return a, b, s


def ntt_inputs():
    rng = rs(102)
    x_i32 = torus32(rng, (3, N))
    x_i32[2, :] = 0
    x_i32[2, 0] = -2**31
    x_i32[2, 1] = 2**31 - 1
    x_i32[2, 1023] = -1
    x_u64 = ff_numbers(rng, (2, N))
    return x_i32, x_u64


def shift_inputs():
    rng = rs(103)
    B = 9
    src = torus32(rng, (B, 2, N))
    powers = numpy.array([0, 1, 1023, 1024, 1025, 2047, 2048 - 1, 517, 1500], numpy.int32)
    bara = torus32(rng, (B, 5), 0, 2 * N)
    bara[:, 3] = powers
    return src, powers, bara


def modswitch_inputs():
    rng = rs(104)
    x = torus32(rng, (64,))
    x[:8] = [0, -1, 2**31 - 1, -2**31, 2**20, 2**20 - 1, -2**20, -2**20 - 1]
    return x


def tgsw_inputs():
    rng = rs(105)
    B = 3
    accum_small = torus32(rng, (B, 2, N), -1000, 1000)      # test_tgsw.py:139-140
    accum_full = torus32(rng, (B, 2, N))
    tr_sample = ff_numbers(rng, (B, 2, 2, N))               # test_tgsw.py:98
    bk = ff_numbers(rng, (3, 2, 2, 2, N))                   # test_tgsw.py:99, bk_len reduced to 3
    return accum_small, accum_full, tr_sample, bk


def keyswitch_inputs():
    rng = rs(106)
    B = 3
    ks_a = torus32(rng, (N, 8, 4, 500), -1000, 1000)        # test_lwe.py:60-62
    ks_b = torus32(rng, (N, 8, 4), -1000, 1000)
    ks_cv = rng.uniform(-1, 1, size=(N, 8, 4)).astype(numpy.float32)
    ks_a[:, :, 0, :] = 0
    ks_b[:, :, 0] = 0
    ks_cv[:, :, 0] = 0
    src_a = torus32(rng, (B, N))
    src_b = torus32(rng, (B,), -1000, 1000)
    return ks_a, ks_b, ks_cv, src_a, src_b


def extract_inputs():
    rng = rs(107)
    return torus32(rng, (4, 2, N))


def linear_inputs():
    rng = rs(108)
    B = 5
    return ((torus32(rng, (B, 500)), torus32(rng, (B,))), (torus32(rng, (B, 500)), torus32(rng, (B,))))


GATE_SEED = 20260923   # the seed SURVEY.md Appendix D validated end to end
GATE_BITS_A = [True, True, False, False]
GATE_BITS_B = [True, False, True, False]
GATE_BITS_C = [False, True, False, True]
