"""Kernel-level parity: CUDA (through the C ABI) == CPU oracle == reference golden vectors, bit for bit.
Mirrors the reference's per-kernel tests (test/test_transform/test_computation.py, test_arithmetic.py,
test_tgsw.py, test_lwe.py, test_tlwe.py)."""
import numpy
import pytest
import torch

import gen_inputs as G
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng():
    from nufhe_b200.engine import Engine
    return Engine()


def dev_u64(eng, arr):
    return eng.to_device(numpy.ascontiguousarray(arr, numpy.uint64))


def test_arithmetic(eng, golden):
    from nufhe_b200 import _native as nv
    g = golden('arithmetic')
    a, b, s = G.arithmetic_inputs()
    da, db = dev_u64(eng, a), dev_u64(eng, b)
    for op, key in ((nv.FF_MUL, 'mul'), (nv.FF_ADD, 'add'), (nv.FF_SUB, 'sub'), (nv.FF_MUL_PREPARED, 'mul_prepared')):
        assert (eng.to_host(eng.ff_op(op, da, db), True) == g[key]).all(), key
    assert (eng.to_host(eng.ff_op(nv.FF_PREPARE, da), True) == g['prepare_for_mul']).all()
    ds = dev_u64(eng, s.astype(numpy.uint64))
    assert (eng.to_host(eng.ff_op(nv.FF_LSH, da, ds), True) == g['lsh']).all()
    assert (eng.to_host(eng.ff_op(nv.FF_LSH_CONST, da, ds), True) == g['lsh']).all()


def test_constant_shift_paths_exhaustive(eng):
    """Every compile-time shift 0..191 used by the transforms, on edge values (incl. the non-canonical
    representative p of zero that the device arithmetic tolerates) and random field elements."""
    from nufhe_b200 import _native as nv
    rng = G.rs(21)
    edge = numpy.concatenate([G.FF_EDGE, numpy.array([G.P, 2**64 - 1, 2**63 - 1, 0xffffffff00000000,
                                                      0xfffffffeffffffff, 0x00000000ffffffff], numpy.uint64)])
    vals = numpy.concatenate([edge, G.ff_numbers(rng, (256 - edge.size,))])
    a = numpy.repeat(vals, 192)
    s = numpy.tile(numpy.arange(192, dtype=numpy.uint64), vals.size)
    got = eng.to_host(eng.ff_op(nv.FF_LSH_CONST, dev_u64(eng, a), dev_u64(eng, s)), True)
    want = O.ff_lsh(a, s.astype(numpy.uint32))
    assert (got == want).all()


def test_arithmetic_random_vs_oracle(eng):
    from nufhe_b200 import _native as nv
    rng = G.rs(11)
    a, b = G.ff_numbers(rng, (1 << 16,)), G.ff_numbers(rng, (1 << 16,))
    da, db = dev_u64(eng, a), dev_u64(eng, b)
    assert (eng.to_host(eng.ff_op(nv.FF_MUL, da, db), True) == O.ff_mul(a, b)).all()
    assert (eng.to_host(eng.ff_op(nv.FF_ADD, da, db), True) == O.ff_add(a, b)).all()
    assert (eng.to_host(eng.ff_op(nv.FF_SUB, da, db), True) == O.ff_sub(a, b)).all()
    s = rng.randint(0, 192, size=a.shape).astype(numpy.uint64)
    assert (eng.to_host(eng.ff_op(nv.FF_LSH, da, dev_u64(eng, s)), True) == O.ff_lsh(a, s.astype(numpy.uint32))).all()


def test_ntt_golden(eng, golden):
    g = golden('ntt')
    x_i32, x_u64 = G.ntt_inputs()
    assert (eng.to_host(eng.ntt_forward_i32(eng.to_device(x_i32)), True) == g['fwd_i32']).all()
    assert (eng.to_host(eng.ntt_forward_u64(dev_u64(eng, x_u64)), True) == g['fwd_u64']).all()
    assert (eng.to_host(eng.ntt_inverse_u64(dev_u64(eng, x_u64)), True) == g['inv_u64']).all()
    assert (eng.to_host(eng.ntt_inverse_i32(dev_u64(eng, x_u64))) == g['inv_i32']).all()


@pytest.mark.parametrize('batch', [1, 3, 257, 5000])
def test_ntt_vs_oracle(eng, batch):
    rng = G.rs(200 + batch)
    x = G.torus32(rng, (batch, 1024))
    f = eng.ntt_forward_i32(eng.to_device(x))
    ref = O.ntt_forward_i32(x)
    assert (eng.to_host(f, True) == ref).all()
    assert (eng.to_host(eng.ntt_inverse_i32(f)) == x).all()
    y = G.ff_numbers(rng, (batch, 1024))
    assert (eng.to_host(eng.ntt_inverse_u64(dev_u64(eng, y)), True) == O.ntt_inverse_u64(y)).all()


def test_ntt_empty_batch(eng):
    out = eng.ntt_forward_i32(eng.empty((0, 1024), torch.int32))
    assert out.shape == (0, 1024)


def test_ntt_convolution_property(eng):
    # NTT -> pointwise product -> INTT == negacyclic product mod 2^32 (test_computation.py:71-124)
    from nufhe_b200 import _native as nv
    rng = G.rs(12)
    a = G.torus32(rng, (8, 1024))
    b = G.torus32(rng, (8, 1024), -1000, 1000)
    fa, fb = eng.ntt_forward_i32(eng.to_device(a)), eng.ntt_forward_i32(eng.to_device(b))
    prod = eng.to_host(eng.ntt_inverse_i32(eng.ff_op(nv.FF_MUL, fa, fb)))
    assert (prod == O.poly_mul_i32(b[0], a)[0:1]).all() or True   # shape smoke; exact check below
    for q in range(8):
        full = numpy.convolve(a[q].astype(object), b[q].astype(object))
        neg = full[:1024].copy()
        neg[:1023] -= full[1024:]
        want = numpy.array([int(v) % 2**32 for v in neg], numpy.uint64).astype(numpy.uint32)
        assert (prod[q].view(numpy.uint32) == want).all()


def test_external_product_golden(eng, golden):
    g = golden('tgsw')
    accum_small, accum_full, tr_sample, bk = G.tgsw_inputs()
    bk_int = eng.bk_prepare(dev_u64(eng, bk))
    acc = eng.to_device(accum_small)
    assert (eng.to_host(eng.external_product(acc, bk_int, 2)) == g['ext_small']).all()
    acc = eng.to_device(accum_full)
    assert (eng.to_host(eng.external_product(acc, bk_int, 0)) == g['ext_full']).all()


@pytest.mark.parametrize('batch', [1, 4, 5, 67])
def test_external_product_vs_oracle(eng, batch):
    rng = G.rs(300 + batch)
    bk = G.ff_numbers(rng, (2, 2, 2, 2, 1024))
    accum = G.torus32(rng, (batch, 2, 1024))
    bk_int = eng.bk_prepare(dev_u64(eng, bk))
    got = eng.to_host(eng.external_product(eng.to_device(accum), bk_int, 1))
    assert (got == O.tgsw_external_mul(accum, bk, 1)).all()


def test_blind_rotate_explicit_vs_oracle(eng):
    # BlindRotate_gpu semantics with an explicit accumulator / bara and a short random-field key
    rng = G.rs(13)
    n, B = 6, 5
    bk = G.ff_numbers(rng, (n, 2, 2, 2, 1024))
    accum = G.torus32(rng, (B, 2, 1024))
    bara = G.torus32(rng, (B, n), 0, 2048)
    bara[0, :4] = [0, 1024, 1023, 2047]
    bk_int = eng.bk_prepare(dev_u64(eng, bk))
    out_a, out_b, acc_out = eng.blind_rotate(eng.to_device(accum), eng.to_device(bara), bk_int, return_accum=True)
    want = O.blind_rotate(accum, bk, bara)
    assert (eng.to_host(acc_out) == want).all()
    ea, eb = O.tlwe_extract_lwe_samples(want)
    assert (eng.to_host(out_a) == ea).all() and (eng.to_host(out_b) == eb).all()


def test_keyswitch_golden(eng, golden):
    g = golden('keyswitch')
    ks_a, ks_b, ks_cv, src_a, src_b = G.keyswitch_inputs()
    ks = (eng.to_device(ks_a), eng.to_device(ks_b), eng.to_device(ks_cv))
    ra, rb, rcv = eng.keyswitch(ks, (eng.to_device(src_a), eng.to_device(src_b)), want_cv=True)
    assert (eng.to_host(ra) == g['res_a']).all() and (eng.to_host(rb) == g['res_b']).all()
    assert numpy.allclose(eng.to_host(rcv), g['res_cv'], rtol=1e-4, atol=1e-4)


def test_keyswitch_ragged_batches(eng):
    ks_a, ks_b, ks_cv, _, _ = G.keyswitch_inputs()
    ks = (eng.to_device(ks_a), eng.to_device(ks_b), eng.to_device(ks_cv))
    rng = G.rs(14)
    for B in (1, 7, 8, 9, 33, 160, 600):
        src_a, src_b = G.torus32(rng, (B, 1024)), G.torus32(rng, (B,))
        src2_a, src2_b = G.torus32(rng, (B, 1024)), G.torus32(rng, (B,))
        ra, rb, _ = eng.keyswitch(ks, (eng.to_device(src_a), eng.to_device(src_b)))
        wa, wb, _ = O.lwe_keyswitch(ks_a, ks_b, ks_cv, src_a, src_b)
        assert (eng.to_host(ra) == wa).all() and (eng.to_host(rb) == wb).all()
        # fused (0,c) + src1 + src2 prologue used by gate_mux
        ra, rb, _ = eng.keyswitch(ks, (eng.to_device(src_a), eng.to_device(src_b)),
                                  (eng.to_device(src2_a), eng.to_device(src2_b)), c=2**29)
        with numpy.errstate(over='ignore'):
            sa = (src_a + src2_a).astype(numpy.int32)
            sb = (src_b + src2_b + numpy.int32(2**29)).astype(numpy.int32)
        wa, wb, _ = O.lwe_keyswitch(ks_a, ks_b, ks_cv, sa, sb)
        assert (eng.to_host(ra) == wa).all() and (eng.to_host(rb) == wb).all()


def test_lwe_affine(eng, golden):
    g = golden('small')
    a, b = G.linear_inputs()
    da = (eng.to_device(a[0]), eng.to_device(a[1]))
    db = (eng.to_device(b[0]), eng.to_device(b[1]))
    for name in ('nand', 'xor', 'andny'):
        num, den, sa, sb = O.GATE_TABLE[name]
        res = (torch.empty_like(da[0]), torch.empty_like(da[1]))
        eng.lwe_affine(res, da, db, O.phase_to_t32(num, den), sa, sb)
        assert (eng.to_host(res[0]) == g['lin_%s_a' % name]).all()
        assert (eng.to_host(res[1]) == g['lin_%s_b' % name]).all()


def test_multi_kernel_steps_golden(eng, golden):
    """The separate steps of the reference's multi-kernel bootstrap (bootstrap.py:96-196): mod-switch, the three
    polynomial rotation modes, trivial sample, sample extraction, accumulator addition -- against the reference's
    own closures (goldens) and the oracle."""
    g = golden('small')
    x = G.modswitch_inputs()
    out = eng.t32_to_phase(eng.empty(x.shape, torch.int32), eng.to_device(x), 2048)
    assert (eng.to_host(out) == g['phase']).all()
    src, powers, bara = G.shift_inputs()
    dsrc = eng.to_device(src)
    for mode, key in ((eng.SHIFT_PLAIN, 'shift_plain'), (eng.SHIFT_INVERT, 'shift_inverted')):
        res = eng.shift_torus_polynomial(torch.empty_like(dsrc), dsrc, eng.to_device(powers), polys_per_power=2, mode=mode)
        assert (eng.to_host(res) == g[key]).all(), key
    res = eng.shift_torus_polynomial(torch.empty_like(dsrc), dsrc, eng.to_device(bara), power_idx=3, polys_per_power=2,
                                     mode=eng.SHIFT_MINUS_ONE)
    assert (eng.to_host(res) == g['shift_minus_one']).all()
    acc = eng.to_device(G.extract_inputs())
    ea, eb = eng.empty((4, 1024), torch.int32), eng.empty((4,), torch.int32)
    eng.tlwe_extract_lwe_samples(ea, eb, acc)
    assert (eng.to_host(ea) == g['extract_a']).all() and (eng.to_host(eb) == g['extract_b']).all()
    triv = eng.empty((9, 2, 1024), torch.int32)
    cv = torch.full((9,), 3.0, dtype=torch.float32, device=triv.device)
    eng.tlwe_noiseless_trivial(triv, cv, eng.to_device(numpy.ascontiguousarray(src[:, 0, :])))
    assert (eng.to_host(triv) == g['trivial']).all() and (eng.to_host(cv) == 0).all()


@pytest.mark.parametrize('batch', [1, 33])
def test_multi_kernel_steps_vs_oracle(eng, batch):
    rng = G.rs(300 + batch)
    src = G.torus32(rng, (batch, 2, 1024))
    bara = G.torus32(rng, (batch, 7), 0, 2048)
    dsrc = eng.to_device(src)
    for idx in (0, 6):
        res = eng.shift_torus_polynomial(torch.empty_like(dsrc), dsrc, eng.to_device(bara), power_idx=idx,
                                         polys_per_power=2, mode=eng.SHIFT_MINUS_ONE)
        assert (eng.to_host(res) == O.shift_torus_polynomial(src, bara, idx, minus_one=True)).all()
    one = numpy.ascontiguousarray(src[:, :1, :])
    pw = numpy.ascontiguousarray(bara[:, 0])
    res = eng.shift_torus_polynomial(eng.empty(one.shape, torch.int32), eng.to_device(one), eng.to_device(pw),
                                     mode=eng.SHIFT_INVERT)
    assert (eng.to_host(res) == O.shift_torus_polynomial(one, pw, invert_powers=True)).all()
    ea, eb = eng.empty((batch, 1024), torch.int32), eng.empty((batch,), torch.int32)
    eng.tlwe_extract_lwe_samples(ea, eb, dsrc)
    oa, ob = O.tlwe_extract_lwe_samples(src)
    assert (eng.to_host(ea) == oa).all() and (eng.to_host(eb) == ob).all()
    other = G.torus32(rng, (batch, 2, 1024))
    acc = eng.to_device(src.copy())
    cv1 = torch.full((batch,), 0.25, dtype=torch.float32, device=acc.device)
    cv2 = torch.full((batch,), 0.5, dtype=torch.float32, device=acc.device)
    eng.tlwe_add_to(acc, eng.to_device(other), cv1, cv2)
    want = (src.view(numpy.uint32) + other.view(numpy.uint32)).view(numpy.int32)
    assert (eng.to_host(acc) == want).all() and (eng.to_host(cv1) == 0.75).all()
    x = G.torus32(rng, (batch, 500))
    out = eng.t32_to_phase(eng.empty(x.shape, torch.int32), eng.to_device(x), 2048)
    assert (eng.to_host(out) == O.t32_to_phase(x, 2048)).all()


def test_transform_interface(eng):
    """nufhe's `Transform` / `ForwardTransform` / `InverseTransform` (transform/computation.py:28-99,
    polynomial_transform_ntt.py:120-131) on top of the stand-alone kernels."""
    from nufhe_b200.transform import Transform, ForwardTransform, InverseTransform, get_transform
    rng = G.rs(41)
    x = G.torus32(rng, (3, 5, 1024))
    fwd = ForwardTransform((3, 5), 1024, None).compile(eng)
    inv = InverseTransform((3, 5), 1024, None).compile(eng)
    tr = eng.empty((3, 5, 1024), torch.int64)
    fwd(tr, eng.to_device(x))
    assert (eng.to_host(tr, True) == O.ntt_forward_i32(x.reshape(15, 1024)).reshape(3, 5, 1024)).all()
    back = eng.empty((3, 5, 1024), torch.int32)
    inv(back, tr)
    assert (eng.to_host(back) == x).all()
    ff = G.ff_numbers(rng, (4, 1024))
    t2 = Transform(None, (4,), kernel_repetitions=2).compile(eng)
    out = eng.empty((4, 1024), torch.int64)
    t2(out, dev_u64(eng, ff))
    assert (eng.to_host(out, True) == O.ntt_forward_u64(ff)).all()
    with pytest.raises(ValueError):
        t2(out, dev_u64(eng, ff[:2]))
    with pytest.raises(ValueError):
        get_transform('FFT')


def test_external_product_steps_golden(eng, golden):
    """The external product as separate steps (decompose, forward transforms, MAC on the reference's key layout,
    inverse transforms) against the reference's closures: k = 1 (tests/golden/tgsw.npz) and k = 2
    (tests/golden/k2_small.npz, make_golden_k2.py)."""
    g = golden('tgsw')
    accum_small, accum_full, tr_sample, bk = G.tgsw_inputs()
    offset = -2145386496                                                 # TGswParams.offset, SURVEY 8 a8
    dec = eng.tgsw_decompose(eng.to_device(accum_full), 2, 10, offset)
    assert (eng.to_host(dec) == g['decomp']).all()
    mac = eng.tgsw_mac(dev_u64(eng, tr_sample), dev_u64(eng, bk[1]), 1, 2)
    assert (eng.to_host(mac, True) == g['mac']).all()
    for acc, row, key in ((accum_small, 2, 'ext_small'), (accum_full, 0, 'ext_full')):
        d = eng.tgsw_decompose(eng.to_device(acc), 2, 10, offset)
        res = eng.ntt_inverse_i32(eng.tgsw_mac(eng.ntt_forward_i32(d), dev_u64(eng, bk[row]), 1, 2))
        assert (eng.to_host(res) == g[key]).all(), key
    # k = 2
    g2 = golden('k2_small')
    rng = G.rs(205)
    accum = G.torus32(rng, (2, 3, 1024))
    tr2 = G.ff_numbers(rng, (2, 3, 2, 1024))
    bk2 = G.ff_numbers(rng, (2, 3, 2, 3, 1024))
    dec = eng.tgsw_decompose(eng.to_device(accum), 2, 10, offset)
    assert (eng.to_host(dec) == g2['decomp']).all()
    mac = eng.tgsw_mac(dev_u64(eng, tr2), dev_u64(eng, bk2[1]), 2, 2)
    assert (eng.to_host(mac, True) == g2['mac']).all()
    res = eng.ntt_inverse_i32(eng.tgsw_mac(eng.ntt_forward_i32(dec), dev_u64(eng, bk2[0]), 2, 2))
    assert (eng.to_host(res) == g2['ext']).all()


@pytest.mark.parametrize('k,batch', [(1, 5), (2, 7), (3, 2)])
def test_external_product_steps_vs_oracle(eng, k, batch):
    rng = G.rs(600 + k)
    accum = G.torus32(rng, (batch, k + 1, 1024))
    bk_row = G.ff_numbers(rng, (k + 1, 2, k + 1, 1024))
    dec = eng.tgsw_decompose(eng.to_device(accum), 2, 10, -2145386496)
    assert (eng.to_host(dec) == O.tgsw_decompose_k(accum)).all()
    tr = eng.ntt_forward_i32(dec)
    mac = eng.tgsw_mac(tr, dev_u64(eng, bk_row), k, 2)
    assert (eng.to_host(mac, True) == O.tgsw_mac_k(eng.to_host(tr, True), bk_row)).all()
    assert (eng.to_host(eng.ntt_inverse_i32(mac)) == O.tgsw_external_mul_k(accum, bk_row)).all()


def test_keyswitch_variances_are_reproducible_across_launch_shapes(eng):
    """The variance is a float32 sum of fixed shape (block sums of 8 coefficients, then the blocks in order), whatever
    the launch: batches small enough to split the input coefficients over many CTAs (1 ciphertext: 128 slices) give
    the same bits as a big batch that does not split, run after run; and it agrees with the oracle's plain running
    sum to float accuracy (the reference's own test uses allclose, test_lwe.py:101)."""
    rng = G.rs(15)
    ks_a, ks_b, _, _, _ = G.keyswitch_inputs()
    ks_cv = (rng.uniform(0.0, 1e-6, size=ks_b.shape) * (numpy.arange(4) > 0)).astype(numpy.float32)   # row d = 0 is padding
    ks = (eng.to_device(ks_a), eng.to_device(ks_b), eng.to_device(ks_cv))
    B = 4800                                                       # 150 tiles >= 148 SMs: no split
    src_a, src_b = G.torus32(rng, (B, 1024)), G.torus32(rng, (B,))
    da, db = eng.to_device(src_a), eng.to_device(src_b)
    _, _, cv_big = eng.keyswitch(ks, (da, db), want_cv=True)
    cv_big = eng.to_host(cv_big)
    _, _, want = O.lwe_keyswitch(ks_a, ks_b, ks_cv, src_a[:70], src_b[:70])
    assert numpy.allclose(cv_big[:70], want, rtol=1e-5)
    for b in (1, 5, 33, 70):
        for _ in range(2):
            ra, rb, cv = eng.keyswitch(ks, (da[:b].contiguous(), db[:b].contiguous()), want_cv=True)
            assert (eng.to_host(cv) == cv_big[:b]).all(), b


def test_tlwe_noiseless_trivial_writes_one_variance_per_sample(eng):
    """Regression: the kernel used to zero B * (k + 1) floats of a (B,) variance array.  A canary right behind the
    variances (same allocation) must survive, for k = 1 and k = 2 and batches beyond the allocator's rounding."""
    for k, B in ((1, 300), (2, 129)):
        buf = torch.full((2 * B + 64,), 7.0, dtype=torch.float32, device=eng.device)
        cv = buf[:B]
        acc = eng.empty((B, k + 1, 1024), torch.int32)
        mu = torch.arange(B * 1024, dtype=torch.int32, device=eng.device).reshape(B, 1024)
        eng.tlwe_noiseless_trivial(acc, cv, mu)
        assert bool((buf[:B] == 0).all()) and bool((buf[B:] == 7.0).all())
        assert bool((acc[:, k] == mu).all()) and bool((acc[:, :k] == 0).all())


@pytest.mark.parametrize('n', [500, 501, 1024, 3])
def test_lwe_dot(eng, n):
    """nb_lwe_dot (LweEncrypt / LweDecrypt, lwe_gpu.mako:205-262): wrap-around dot product with the key plus addends,
    16-byte path (n % 4 == 0) and scalar path, against exact integer arithmetic."""
    rng = G.rs(600 + n)
    for B in (1, 37, 2000):
        a = G.torus32(rng, (B, n))
        key = rng.randint(0, 2, size=(n,)).astype(numpy.int32)
        add1, add2 = G.torus32(rng, (B,)), G.torus32(rng, (B,))
        dot = (a.astype(numpy.int64) * key.astype(numpy.int64)).sum(-1)

        def wrap(x):
            return ((x + 2**31) % 2**32 - 2**31).astype(numpy.int32)
        da, dk = eng.to_device(a), eng.to_device(key)
        got = eng.to_host(eng.lwe_dot(da, dk, add1=eng.to_device(add1), add2=eng.to_device(add2), sign=1))
        assert (got == wrap(dot + add1.astype(numpy.int64) + add2)).all()
        got = eng.to_host(eng.lwe_dot(da, dk, add1=eng.to_device(add1), sign=-1))
        assert (got == wrap(add1.astype(numpy.int64) - dot)).all()
        assert (eng.to_host(eng.lwe_dot(da, dk)) == wrap(dot)).all()
    # a general (non-binary) key wraps the same way
    key = G.torus32(rng, (n,))
    a = G.torus32(rng, (5, n))
    want = numpy.array([sum(int(x) * int(k) for x, k in zip(row, key)) for row in a])
    assert (eng.to_host(eng.lwe_dot(eng.to_device(a), eng.to_device(key))).astype(numpy.int64) == (want + 2**31) % 2**32 - 2**31).all()


def test_make_keyswitch_key_small(eng):
    """nb_make_keyswitch_key (lwe_gpu.mako:18-56) on a small shape against the formula of lwe_cpu.py:26-59; the full-size
    key is pinned by the key digests of the reference (test_gpu_api.py::test_seeded_keys_equal_reference_keys)."""
    rng = G.rs(77)
    in_size, t, log2_base, n = 6, 3, 2, 10
    base = 1 << log2_base
    in_key = rng.randint(0, 2, size=(in_size,)).astype(numpy.int32)
    out_key = rng.randint(0, 2, size=(n,)).astype(numpy.int32)
    na, nb_ = G.torus32(rng, (in_size, t, base - 1, n)), G.torus32(rng, (in_size, t, base - 1))
    ks_a = torch.full((in_size, t, base, n), 5, dtype=torch.int32, device=eng.device)
    ks_b = torch.full((in_size, t, base), 5, dtype=torch.int32, device=eng.device)
    ks_cv = torch.full((in_size, t, base), 5.0, dtype=torch.float32, device=eng.device)
    eng.make_keyswitch_key(ks_a, ks_b, ks_cv, eng.to_device(in_key), eng.to_device(out_key), eng.to_device(na),
                           eng.to_device(nb_), log2_base, 0.25)
    a, b, cv = eng.to_host(ks_a), eng.to_host(ks_b), eng.to_host(ks_cv)
    assert (a[:, :, 0] == 0).all() and (b[:, :, 0] == 0).all() and (cv[:, :, 0] == 0).all()
    assert (a[:, :, 1:] == na).all() and (cv[:, :, 1:] == 0.25).all()
    for i in range(in_size):
        for j in range(t):
            for h in range(1, base):
                want = int(in_key[i]) * h * 2 ** (32 - (j + 1) * log2_base) + int(nb_[i, j, h - 1]) + int(
                    (na[i, j, h - 1].astype(numpy.int64) * out_key).sum())
                assert int(b[i, j, h]) == (want + 2**31) % 2**32 - 2**31
