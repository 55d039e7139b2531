"""CPU-only checks of the inline-PTX device arithmetic of csrc/ff.cuh (executed by tests/ptx_emul.py from the
source text) against exact integer arithmetic mod p = 2^64 - 2^32 + 1, on edge values and random inputs.

The Python functions below mirror the few lines of C glue around each asm block (they must be kept in step with
ff.cuh; the asm text itself is read from the file).  Invariant under test: every device primitive maps [0, p]
("almost canonical") inputs to [0, p] outputs congruent to the exact result; ff_sub additionally accepts any
64-bit minuend.
"""
import os
import random

import pytest

from ptx_emul import Machine, extract_asm_blocks, preprocess, M32, M64

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RAW = open(os.path.join(ROOT, 'nufhe_b200', 'csrc', 'ff.cuh')).read()

SRC = preprocess(RAW, {})
P = (1 << 64) - (1 << 32) + 1
EPS = (1 << 32) - 1

HELPERS = {
    'lo32': lambda x: x & M32,
    'hi32': lambda x: (x >> 32) & M32,
    'pack': lambda lo, hi: ((hi & M32) << 32) | (lo & M32),
    'nb_c_eps': EPS,
    'nb_c_pow2': [1 << i for i in range(32)],
}
_BLOCKS = {}


def blocks(fn):
    if fn not in _BLOCKS:
        _BLOCKS[fn] = extract_asm_blocks(SRC, fn)
        assert _BLOCKS[fn], 'no asm block found in %s' % fn
    return _BLOCKS[fn]


def run(fn, index=0, **variables):
    env = dict(HELPERS)
    env.update(variables)
    return Machine().run(blocks(fn)[index], env)


# ---- mirrors of the C glue ------------------------------------------------------------------------
def ff_sub(a, b, chain=False):
    # ff_sub_dev<CHAIN>: block 0 = second borrow chain, block 1 = add-chain fold (the one the build uses)
    e = run('ff_sub_dev', 0 if chain else 1, a=a, b=b)
    return HELPERS['pack'](e['l'], e['h'])


def ff_add(a, b):
    return ff_sub(a, (P - b) & M64)


def ff_add_nc(a, b):
    e = run('ff_add_nc', a=a, b=b)
    return HELPERS['pack'](e['l'], e['h'])


def ff_add_keps(v0, v1, k):
    e = run('ff_add_keps', v0=v0, v1=v1, k=k)
    return HELPERS['pack'](e['v0'], (e['v1'] + k) & M32)


def ff_canon_dev(v0, v1):
    e = run('ff_canon_dev', v0=v0, v1=v1)
    return ff_add_keps(v0, v1, e['f'])


def ff_reduce_limbs_nc(l, m, h0, h1):
    e = run('ff_reduce_limbs_nc', l=l, m=m, h0=h0, h1=h1)
    return e['r0'] | (e['r1'] << 32)


def ff_reduce_limbs(l, m, h0, h1):
    v = ff_reduce_limbs_nc(l, m, h0, h1)
    return ff_canon_dev(v & M32, v >> 32)


def mul128(a, b):
    e = run('mul128', a=a, b=b)
    return e['r0'], e['r1'], e['r2'], e['r3']


def ff_mul(a, b):
    return ff_reduce_limbs(*mul128(a, b))


def ff_dot4(a, b):
    c = list(mul128(a[0], b[0])) + [0]
    for k in range(1, 4):
        e = run('mac128', a=a[k], b=b[k], c0=c[0], c1=c[1], c2=c[2], c3=c[3], c4=c[4])
        c = [e['c%d' % i] for i in range(5)]
    return ff_sub(ff_reduce_limbs(*c[:4]), (c[4] << 32) & M64)


def mulwide(a, b):
    e = run('mulwide', a=a, b=b)
    return e['t'] & M32, e['t'] >> 32


def ff_comb_a(y0, y1, y2):
    e = run('ff_comb_a', y0=y0, y1=y1, y2=y2)
    return ff_add_keps(e['r0'], e['r1'], e['k'])


def ff_comb_b(y0, y1, y2):
    e = run('ff_comb_b', y0=y0, y1=y1, y2=y2)
    return HELPERS['pack'](e['r0'], e['r1'])


def ff_comb_c(y0, y1, y2, neg=False):
    u0, u1 = mulwide(y0, EPS)
    u, v = HELPERS['pack'](u0, u1), HELPERS['pack'](y1, y2)
    return ff_sub(v, u) if neg else ff_sub(u, v)


def ff_shl_dev(x, S):
    s = S % 192
    s96, negate = s % 96, s >= 96
    r, q = s96 % 32, s96 // 32
    if r == 0:
        y0, y1, y2 = x & M32, x >> 32, 0
    else:
        y0, y1, y2 = (x << r) & M32, (x >> (32 - r)) & M32, x >> (64 - r)
    if q == 0:
        v = ff_comb_a(y0, y1, y2)
        return (P - v) if negate else v
    if q == 1:
        v = ff_comb_b(y0, y1, y2)
        return (P - v) if negate else v
    return ff_comb_c(y0, y1, y2, negate)


def ff_shl(x, S):
    s = S % 192
    if s == 0:
        return x
    if s == 96:
        return P - x
    return ff_shl_dev(x, S)


# ---- inputs -------------------------------------------------------------------------------------
EDGE64 = sorted({0, 1, 2, 3, (1 << 31) - 1, 1 << 31, EPS - 1, EPS, EPS + 1, EPS + 2, 1 << 33, (1 << 33) + 1,
                 (1 << 63) - 1, 1 << 63, (1 << 63) + EPS, (1 << 64) - (1 << 33), P - (1 << 33) - 1, P - (1 << 33),
                 P - EPS - 2, P - EPS - 1, P - EPS, P - EPS + 1, P - 3, P - 2, P - 1, P,
                 0x5555555555555555, 0xaaaaaaaaaaaaaaaa})
LOOSE64 = [P + 1, P + 2, (1 << 64) - EPS, (1 << 64) - 2, (1 << 64) - 1]
EDGE32 = [0, 1, 2, 3, (1 << 31) - 1, 1 << 31, (1 << 31) + 1, M32 - 2, M32 - 1, M32]
RNG = random.Random(20260923)


def rand_field(n):
    return [RNG.randrange(P + 1) for _ in range(n)]


def in_range(v):
    return 0 <= v <= P


def test_parser_sees_every_device_sequence():
    for fn, n in (('ff_sub_dev', 2), ('mul128', 1), ('mac128', 1), ('ff_add_keps', 1), ('ff_canon_dev', 1),
                  ('ff_reduce_limbs_nc', 1), ('ff_comb_a', 1), ('ff_comb_b', 1), ('mulwide', 1)):
        assert len(blocks(fn)) == n, fn


def test_carry_flag_families_are_never_mixed():
    # the interpreter raises if a subtract reads an add's flag (or the reverse); run every block once
    ff_dot4([1, 2, 3, 4], [5, 6, 7, 8])
    ff_shl(12345, 37), ff_shl(12345, 70), ff_shl(12345, 5)


def test_sub_add_all_edge_pairs():
    for chain in (False, True):
        for a in EDGE64 + LOOSE64:
            for b in EDGE64:
                got = ff_sub(a, b, chain)
                want = a - b if a >= b else a - b + P          # exact, no further reduction (a may be loose)
                assert got == want, (chain, hex(a), hex(b), hex(got), hex(want))
                if a <= P:
                    assert in_range(got)
    for a in EDGE64:
        for b in EDGE64:
            got = ff_add(a, b)
            assert in_range(got) and got % P == (a + b) % P, (hex(a), hex(b), hex(got))
    for a, b in zip(rand_field(2000), rand_field(2000)):
        assert ff_sub(a, b) % P == (a - b) % P and in_range(ff_sub(a, b))
        assert ff_add(a, b) % P == (a + b) % P and in_range(ff_add(a, b))


def test_add_nc_all_edge_pairs():
    """ff_add_nc: the sum as any 64-bit value of the right residue; above p only without a wrap, and then its high
    limb is 2^32 - 1 (what the butterfly networks' trigger watches)."""
    assert len(blocks('ff_add_nc')) == 1
    pairs = [(a, b) for a in EDGE64 for b in EDGE64] + list(zip(rand_field(3000), rand_field(3000)))
    for a, b in pairs:
        got = ff_add_nc(a, b)
        assert 0 <= got < 1 << 64 and got % P == (a + b) % P, (hex(a), hex(b), hex(got))
        if got > P:
            assert a + b < 1 << 64 and got >> 32 == M32, (hex(a), hex(b), hex(got))
    # the window (p, 2^64) is reachable: p - 1 + 5 does not wrap and is left alone
    assert ff_add_nc(P - 1, 5) == P + 4


def test_canon_and_keps():
    for v in EDGE64 + LOOSE64 + [RNG.randrange(1 << 64) for _ in range(500)]:
        got = ff_canon_dev(v & M32, v >> 32)
        assert got == (v - P if v > P else v), hex(v)
        for k in (0, 1):
            assert ff_add_keps(v & M32, v >> 32, k) == (v + k * EPS) & M64


def test_reduce_limbs_all_edge_limbs():
    phi = 1 << 32
    for l in EDGE32:
        for m in EDGE32:
            for h0 in EDGE32:
                for h1 in EDGE32:
                    got = ff_reduce_limbs(l, m, h0, h1)
                    want = (l + m * phi + h0 * phi ** 2 + h1 * phi ** 3) % P
                    assert in_range(got) and got % P == want, (l, m, h0, h1, hex(got))
                    # the uncanonicalised form: any 64-bit value of the right residue, and the rare-path trigger
                    # (high limb all ones) fires whenever it lies above p
                    nc = ff_reduce_limbs_nc(l, m, h0, h1)
                    assert 0 <= nc < 1 << 64 and nc % P == want
                    assert nc <= P or nc >> 32 == M32
    for _ in range(3000):
        l, m, h0, h1 = (RNG.randrange(1 << 32) for _ in range(4))
        got = ff_reduce_limbs(l, m, h0, h1)
        assert in_range(got) and got % P == (l + m * phi + h0 * phi ** 2 + h1 * phi ** 3) % P


def test_mul_and_dot4():
    xs = EDGE64
    for a in xs:
        for b in xs:
            r = mul128(a, b)
            assert sum(v << (32 * i) for i, v in enumerate(r)) == a * b
            got = ff_mul(a, b)
            assert in_range(got) and got % P == a * b % P, (hex(a), hex(b))
    worst = [P, P, P, P]
    assert ff_dot4(worst, worst) % P == 0
    top = [(1 << 64) - 1] * 4                               # loose operands: the 130-bit accumulator still holds
    assert ff_dot4(top, top) % P == 4 * top[0] * top[0] % P
    for _ in range(300):
        a, b = rand_field(4), rand_field(4)
        if RNG.random() < 0.3:
            a[RNG.randrange(4)] = RNG.choice(EDGE64)
            b[RNG.randrange(4)] = RNG.choice(EDGE64)
        got = ff_dot4(a, b)
        assert in_range(got) and got % P == sum(x * y for x, y in zip(a, b)) % P


def ff_dotn_sub_nc(a, b, c):
    """ff_dot4_sub_nc / ff_dot2_sub_nc (the MAC of the fused bootstrap, 4 terms; the pair shape's, 2 terms)."""
    acc = list(mul128(a[0], b[0])) + [0]
    for k in range(1, len(a)):
        e = run('mac128', a=a[k], b=b[k], c0=acc[0], c1=acc[1], c2=acc[2], c3=acc[3], c4=acc[4])
        acc = [e['c%d' % i] for i in range(5)]
    return ff_sub(ff_sub(ff_reduce_limbs_nc(*acc[:4]), (acc[4] << 32) & M64), c)


def test_dot_products_minus_correction_keep_the_residue():
    """The not-canonicalised MAC results: right residue, and a value above p always shows a high limb of 2^32 - 1 --
    the trigger of the rare-path fix-up (br_phases.cuh: canon_needed)."""
    for n in (2, 4):
        cases = [([P] * n, [P] * n, 0), ([P - 1] * n, [P - 1] * n, P), ([P] * n, [1] * n, P - 1)]
        for _ in range(400):
            a, b = rand_field(n), rand_field(n)
            if RNG.random() < 0.3:
                a[RNG.randrange(n)] = RNG.choice(EDGE64) % (P + 1)
                b[RNG.randrange(n)] = RNG.choice(EDGE64) % (P + 1)
            cases.append((a, b, RNG.choice([0, 1, P - 1, P, RNG.randrange(P)])))
        for a, b, c in cases:
            got = ff_dotn_sub_nc(a, b, c)
            assert 0 <= got < (1 << 64) and got % P == (sum(x * y for x, y in zip(a, b)) - c) % P
            assert got <= P or (got >> 32) == M32


def test_limb_combinations_on_edge_limbs():
    phi = 1 << 32
    ys2 = [0, 1, 2, (1 << 30), (1 << 31) - 2, (1 << 31) - 1]          # y2 < 2^31 by construction (r <= 31)
    for y0 in EDGE32:
        for y1 in EDGE32:
            for y2 in ys2:
                val = y0 + y1 * phi + y2 * phi ** 2
                for fn, q, neg in ((ff_comb_a, 0, False), (ff_comb_b, 1, False)):
                    got = fn(y0, y1, y2)
                    assert in_range(got) and got % P == val * phi ** q % P, (fn.__name__, y0, y1, y2, hex(got))
                if (y1 | (y2 << 32)) <= P:
                    for neg in (False, True):
                        got = ff_comb_c(y0, y1, y2, neg)
                        want = val * phi ** 2 % P
                        assert in_range(got) and got % P == (P - want if neg else want) % P


@pytest.mark.parametrize('chunk', range(4))
def test_every_constant_shift(chunk):
    xs = EDGE64 + rand_field(12)
    for S in range(chunk * 48, (chunk + 1) * 48):
        for x in xs:
            got = ff_shl(x, S)
            assert in_range(got) and got % P == (x << S) % P, (S, hex(x), hex(got))
