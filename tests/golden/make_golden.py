"""Generate tests/golden/*.npz by running the UNMODIFIED reference's NumPy closures
(/root/reference/nufhe/*_cpu.py, transform/ntt_cpu.py, transform/ntt.py) on the seeded inputs of
gen_inputs.py.  Run once in the build container (the reference is not available on the GPU box):

    python tests/golden/make_golden.py [--skip-gate]

The gate fixture (reference keygen + the full 500-step blind rotation through the reference closures,
composed as nufhe/bootstrap.py:206-229 prescribes) takes ~6 minutes on one core.
"""
import hashlib
import os
import sys
import time

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_inputs as G                      # noqa: E402
from ref_bridge import import_reference     # noqa: E402

nufhe = import_reference()
from nufhe.transform import ntt_cpu                                   # noqa: E402
from nufhe.transform.ntt import ntt_transform_ref                      # noqa: E402
from nufhe.transform.arithmetic import prepare_for_mul_cpu             # noqa: E402
from nufhe import polynomial_transform_ntt as ptn                      # noqa: E402
from nufhe.numeric_functions_cpu import Torus32ToPhaseReference        # noqa: E402
from nufhe.polynomials_cpu import ShiftTorusPolynomialReference       # noqa: E402
from nufhe.tlwe_cpu import (TLweNoiselessTrivialReference, TLweExtractLweSamplesReference,  # noqa: E402
                            TLweEncryptZeroReference)
from nufhe.tgsw_cpu import (tgsw_polynomial_decomp_trf_reference,      # noqa: E402
                            tlwe_transformed_add_mul_to_trf_reference,
                            TGswTransformedExternalMulReference, TGswAddMessageReference)
from nufhe.lwe_cpu import (LweKeyswitchReference, MakeLweKeyswitchKeyReference,  # noqa: E402
                           LweLinearReference, LweNoiselessTrivialConstantReference,
                           LweEncryptReference, LweDecryptReference)
from nufhe.api_low_level import NuFHEParameters                        # noqa: E402
from nufhe.numeric_functions import double_to_t32                      # noqa: E402


def phase_to_t32(phase, mspace_size):
    """nufhe/numeric_functions.py:30-31.  The reference's `Torus32(int)` relied on NumPy-1 wrap-around
    (e.g. phase_to_t32(-1, 8) = 7 * 2^29 -> -2^29) and raises OverflowError on NumPy 2; this is the
    same value with the wrap made explicit."""
    v = (phase % mspace_size) * (2**32 // mspace_size)
    return numpy.int32(v - 2**32 if v >= 2**31 else v)

N = 1024
params = NuFHEParameters(transform_type='NTT')
tgsw_params = params.tgsw_params
tlwe_params = tgsw_params.tlwe_params


def sha(arr):
    return hashlib.sha256(numpy.ascontiguousarray(arr).tobytes()).hexdigest()


def save(name, **kw):
    path = os.path.join(HERE, name + '.npz')
    numpy.savez_compressed(path, **kw)
    print('wrote', path, os.path.getsize(path), 'bytes')


def golden_arithmetic():
    a, b, s = G.arithmetic_inputs()
    ga, gb = ntt_cpu.gnum(a), ntt_cpu.gnum(b)
    save('arithmetic',
         mul=ntt_cpu.gnum_to_u64(ga * gb),
         add=ntt_cpu.gnum_to_u64(ga + gb),
         sub=ntt_cpu.gnum_to_u64(ga - gb),
         mul_prepared=ptn.transformed_space_mul_prepared_ref(a, b),
         prepare_for_mul=prepare_for_mul_cpu(a % numpy.uint64(G.P)),
         lsh=ntt_cpu.gnum_to_u64(ga * numpy.array([ntt_cpu.GaloisNumber(2)**int(x) for x in s])))


def golden_ntt():
    x_i32, x_u64 = G.ntt_inputs()
    save('ntt',
         fwd_i32=ntt_transform_ref(x_i32, i32_conversion=True),
         fwd_u64=ntt_transform_ref(x_u64),
         inv_u64=ntt_transform_ref(x_u64, inverse=True),
         inv_i32=ntt_transform_ref(x_u64, inverse=True, i32_conversion=True))


def golden_small():
    # mod-switch
    x = G.modswitch_inputs()
    ph = numpy.empty(x.shape, numpy.int32)
    Torus32ToPhaseReference(x.shape, 2 * N)(ph, x)
    # rotations, three modes (polynomials.py:83-104)
    src, powers, bara = G.shift_inputs()
    B = src.shape[0]
    r_plain = numpy.empty_like(src)
    ShiftTorusPolynomialReference(N, src.shape[:-1], powers.shape)(r_plain, src, powers, 0)
    r_inv = numpy.empty_like(src)
    ShiftTorusPolynomialReference(N, src.shape[:-1], powers.shape, invert_powers=True)(
        r_inv, src, powers, 0)
    r_m1 = numpy.empty_like(src)
    ShiftTorusPolynomialReference(N, src.shape[:-1], bara.shape, powers_view=True, minus_one=True)(
        r_m1, src, bara, 3)
    # trivial + extract
    acc_in = G.extract_inputs()
    ea = numpy.empty((acc_in.shape[0], N), numpy.int32)
    eb = numpy.empty((acc_in.shape[0],), numpy.int32)
    TLweExtractLweSamplesReference(tlwe_params, (acc_in.shape[0],))(ea, eb, acc_in)
    triv = numpy.empty((B, 2, N), numpy.int32)
    cv = numpy.empty((B,), numpy.float32)
    TLweNoiselessTrivialReference(tlwe_params, (B,))(triv, cv, src[:, 0, :].copy())
    # gate prologue: (0, c) + sa*a + sb*b through the reference's linear closures (gates.py:108-115)
    (a_a, a_b), (b_a, b_b) = G.linear_inputs()
    lin = {}
    for name, (num, den, sa, sb) in dict(nand=(1, 8, -1, -1), xor=(1, 4, 2, 2), andny=(-1, 8, -1, 1)).items():
        t_a = numpy.empty_like(a_a); t_b = numpy.empty_like(a_b); t_cv = numpy.zeros(a_b.shape, numpy.float32)
        LweNoiselessTrivialConstantReference(None)(t_a, t_b, t_cv, phase_to_t32(num, den))
        z = numpy.zeros(a_b.shape, numpy.float32)
        with numpy.errstate(over='ignore'):
            LweLinearReference(None, None, add_result=True)(t_a, t_b, t_cv, a_a, a_b, z, sa)
            LweLinearReference(None, None, add_result=True)(t_a, t_b, t_cv, b_a, b_b, z, sb)
        lin['lin_%s_a' % name] = t_a
        lin['lin_%s_b' % name] = t_b
    save('small', phase=ph, shift_plain=r_plain, shift_inverted=r_inv, shift_minus_one=r_m1,
         extract_a=ea, extract_b=eb, trivial=triv, **lin)


def golden_tgsw():
    accum_small, accum_full, tr_sample, bk = G.tgsw_inputs()
    B = accum_small.shape[0]
    shape = (B,)
    dec = numpy.empty((B, 2, 2, N), numpy.int32)
    tgsw_polynomial_decomp_trf_reference(tgsw_params, shape)(dec, accum_full)
    mac = numpy.empty((B, 2, N), numpy.uint64)
    tlwe_transformed_add_mul_to_trf_reference(tgsw_params, shape, bk.shape[0], None)(mac, tr_sample, bk, 1)
    t = time.time()
    ext_small = accum_small.copy()
    TGswTransformedExternalMulReference(tgsw_params, shape, bk.shape[0], None)(ext_small, bk, 2)
    ext_full = accum_full.copy()
    TGswTransformedExternalMulReference(tgsw_params, shape, bk.shape[0], None)(ext_full, bk, 0)
    print('external mul x2: %.1f s' % (time.time() - t))
    save('tgsw', decomp=dec, mac=mac, ext_small=ext_small, ext_full=ext_full)


def golden_keyswitch():
    ks_a, ks_b, ks_cv, src_a, src_b = G.keyswitch_inputs()
    B = src_b.shape[0]
    ra = numpy.empty((B, 500), numpy.int32)
    rb = numpy.empty((B,), numpy.int32)
    rcv = numpy.empty((B,), numpy.float32)
    with numpy.errstate(over='ignore'):
        LweKeyswitchReference(None, N, 500, 8, 2)(ra, rb, rcv, ks_a, ks_b, ks_cv, src_a, src_b)
    save('keyswitch', res_a=ra, res_b=rb, res_cv=rcv)


def reference_keygen(seed):
    """make_key_pair in the reference's RNG order, every step through the
    reference's own closures."""
    rng = numpy.random.RandomState(seed)
    n = 500
    lwe_key = rng.randint(0, 2, size=(n,), dtype=numpy.int32)
    tlwe_key = rng.randint(0, 2, size=(1, N), dtype=numpy.int32)
    bk_shape = (n, 2, 2)
    noises1 = rng.randint(-2**31, 2**31, size=bk_shape + (1, N), dtype=numpy.int32)
    noises2 = double_to_t32(rng.normal(size=bk_shape + (N,), scale=tlwe_params.min_noise))
    bk = numpy.empty(bk_shape + (2, N), numpy.int32)
    cv = numpy.empty(bk_shape, numpy.float32)
    t = time.time()
    with numpy.errstate(over='ignore'):
        TLweEncryptZeroReference(tlwe_params, bk_shape, tlwe_params.min_noise, None)(
            bk, cv, tlwe_key, noises1, noises2)
        TGswAddMessageReference(tgsw_params, (n,))(bk, lwe_key)
    print('reference BK encrypt: %.1f s' % (time.time() - t)); t = time.time()
    bk_tr = prepare_for_mul_cpu(ntt_transform_ref(bk, i32_conversion=True))
    print('reference BK transform: %.1f s' % (time.time() - t))
    ks_noise = params.in_out_params.min_noise
    noises_b = rng.normal(size=(N, 8, 3), scale=ks_noise)
    noises_b -= noises_b.mean()
    noises_b = double_to_t32(noises_b)
    noises_a = rng.randint(-2**31, 2**31, size=(N, 8, 3, n), dtype=numpy.int32)
    ks_a = numpy.empty((N, 8, 4, n), numpy.int32)
    ks_b = numpy.empty((N, 8, 4), numpy.int32)
    ks_cv = numpy.empty((N, 8, 4), numpy.float32)
    with numpy.errstate(over='ignore'):
        MakeLweKeyswitchKeyReference(N, n, 8, 2, ks_noise)(
            ks_a, ks_b, ks_cv, tlwe_key.ravel(), lwe_key, noises_a, noises_b)
    return rng, lwe_key, tlwe_key, bk, bk_tr, (ks_a, ks_b, ks_cv)


def reference_encrypt(rng, lwe_key, bits):
    bits = numpy.asarray(bits)
    mus = numpy.where(bits, phase_to_t32(1, 8), -phase_to_t32(1, 8)).astype(numpy.int32)
    noise = params.in_out_params.min_noise
    noises_b = double_to_t32(rng.normal(size=bits.shape, scale=noise))
    noises_a = rng.randint(-2**31, 2**31, size=bits.shape + (500,), dtype=numpy.int32)
    a = numpy.empty(bits.shape + (500,), numpy.int32)
    b = numpy.empty(bits.shape, numpy.int32)
    cv = numpy.empty(bits.shape, numpy.float32)
    with numpy.errstate(over='ignore'):
        LweEncryptReference(bits.shape, 500, noise)(a, b, cv, mus, lwe_key, noises_a, noises_b)
    return a, b


def reference_bootstrap(x_a, x_b, bk_tr, ks, mu, no_keyswitch=False):
    """bootstrap(), nufhe/bootstrap.py:206-229 + :154-196 + :96-142 (loop path) from the closures."""
    B = x_b.shape[0]
    n = x_a.shape[-1]
    barb = numpy.empty((B,), numpy.int32)
    bara = numpy.empty((B, n), numpy.int32)
    Torus32ToPhaseReference((B,), 2 * N)(barb, x_b)
    Torus32ToPhaseReference((B, n), 2 * N)(bara, x_a)
    testvect = numpy.full((B, N), mu, numpy.int32)
    testvectbis = numpy.empty((B, N), numpy.int32)
    ShiftTorusPolynomialReference(N, (B,), (B,), invert_powers=True)(testvectbis, testvect, barb, 0)
    acc = numpy.empty((B, 2, N), numpy.int32)
    cv = numpy.empty((B,), numpy.float32)
    TLweNoiselessTrivialReference(tlwe_params, (B,))(acc, cv, testvectbis)
    shift = ShiftTorusPolynomialReference(N, (B, 2), (B, n), powers_view=True, minus_one=True)
    extmul = TGswTransformedExternalMulReference(tgsw_params, (B,), n, None)
    t = time.time()
    for i in range(n):
        tmp = numpy.empty_like(acc)
        with numpy.errstate(over='ignore'):
            shift(tmp, acc, bara, i)
            extmul(tmp, bk_tr, i)
            acc = acc + tmp
        if i % 50 == 0:
            print('  step', i, '%.0f s' % (time.time() - t), flush=True)
    ea = numpy.empty((B, N), numpy.int32)
    eb = numpy.empty((B,), numpy.int32)
    TLweExtractLweSamplesReference(tlwe_params, (B,))(ea, eb, acc)
    if no_keyswitch:
        return ea, eb
    ra = numpy.empty((B, n), numpy.int32)
    rb = numpy.empty((B,), numpy.int32)
    rcv = numpy.empty((B,), numpy.float32)
    with numpy.errstate(over='ignore'):
        LweKeyswitchReference(None, N, n, 8, 2)(ra, rb, rcv, ks[0], ks[1], ks[2], ea, eb)
    return (ra, rb), (ea, eb)


def golden_gate():
    rng, lwe_key, tlwe_key, bk, bk_tr, ks = reference_keygen(G.GATE_SEED)
    c1 = reference_encrypt(rng, lwe_key, G.GATE_BITS_A)
    c2 = reference_encrypt(rng, lwe_key, G.GATE_BITS_B)
    c3 = reference_encrypt(rng, lwe_key, G.GATE_BITS_C)
    # NAND on the first 2 ciphertexts: (0, 1/8) - a - b, gates.py:108-121
    sl = slice(0, 2)
    with numpy.errstate(over='ignore'):
        t_a = (-c1[0][sl] - c2[0][sl]).astype(numpy.int32)
        t_b = (phase_to_t32(1, 8) - c1[1][sl] - c2[1][sl]).astype(numpy.int32)
    (nand_a, nand_b), (ext_a, ext_b) = reference_bootstrap(t_a, t_b, bk_tr, ks, phase_to_t32(1, 8))
    dec = numpy.empty((2,), numpy.int32)
    LweDecryptReference((2,), 500)(dec, nand_a, nand_b, lwe_key)
    print('NAND decrypts to', dec > 0)
    save('gate',
         seed=G.GATE_SEED,
         lwe_key_sha=sha(lwe_key), tlwe_key_sha=sha(tlwe_key), bk_raw_sha=sha(bk), bk_sha=sha(bk_tr),
         ks_a_sha=sha(ks[0]), ks_b_sha=sha(ks[1]),
         bk_row0=bk_tr[0], bk_row499=bk_tr[499],
         c1_a=c1[0], c1_b=c1[1], c2_a=c2[0], c2_b=c2[1], c3_a=c3[0], c3_b=c3[1],
         nand_a=nand_a, nand_b=nand_b, nand_ext_a=ext_a, nand_ext_b=ext_b,
         nand_bits=(dec > 0))


def golden_params_pickle():
    """The parameter object every nufhe dump starts with, pickled by the reference's own class."""
    import pickle
    path = os.path.join(HERE, 'ref_params.pkl')
    with open(path, 'wb') as f:
        pickle.dump(NuFHEParameters(), f)
    print('wrote', path)


if __name__ == '__main__':
    golden_params_pickle()
    golden_arithmetic()
    golden_ntt()
    golden_small()
    golden_tgsw()
    golden_keyswitch()
    if '--skip-gate' not in sys.argv:
        golden_gate()
