"""Golden vectors for `NuFHEParameters(tlwe_mask_size=2)` (the reference supports it on its multi-kernel path only,
blind_rotate.py:53-58), from the reference's own NumPy closures.  Same recipe as make_golden.py: run in the build
container (needs /root/reference), commit the output `tests/golden/k2.npz`.

Contents: gadget decomposition, MAC and one external product with a random field key for k = 2, and one complete
`gate_nand` on two ciphertexts with keys generated in the reference's RNG order (seed `gen_inputs.GATE_SEED`).  The
keys themselves are far too large to commit (the key-switch key is 131 MB), so their SHA-256 digests are stored: the
GPU test regenerates them from the seed with nufhe_b200's own key generation, which therefore has to be bit-identical
too."""
import hashlib
import os
import sys
import time

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_inputs as G                       # noqa: E402
from ref_bridge import import_reference      # noqa: E402

nufhe = import_reference()
from nufhe.api_low_level import NuFHEParameters                                     # noqa: E402
from nufhe.numeric_functions import double_to_t32                                   # noqa: E402
from nufhe.numeric_functions_cpu import Torus32ToPhaseReference                      # noqa: E402
from nufhe.polynomials_cpu import ShiftTorusPolynomialReference                      # noqa: E402
from nufhe.tlwe_cpu import (                                                         # noqa: E402
    TLweNoiselessTrivialReference, TLweExtractLweSamplesReference, TLweEncryptZeroReference)
from nufhe.tgsw_cpu import (                                                         # noqa: E402
    TGswTransformedExternalMulReference, TGswAddMessageReference)
from nufhe.lwe_cpu import (                                                          # noqa: E402
    LweKeyswitchReference, MakeLweKeyswitchKeyReference, LweEncryptReference, LweDecryptReference)
from nufhe.transform.ntt import ntt_transform_ref                                    # noqa: E402
from nufhe.transform.arithmetic import prepare_for_mul_cpu                           # noqa: E402

K = 2
N = 1024
params = NuFHEParameters(transform_type='NTT', tlwe_mask_size=K)
tgsw_params = params.tgsw_params
tlwe_params = tgsw_params.tlwe_params


def phase_to_t32(phase, mspace_size):
    v = (phase % mspace_size) * (2**32 // mspace_size)
    return numpy.int32(v - 2**32 if v >= 2**31 else v)


def sha(arr):
    return hashlib.sha256(numpy.ascontiguousarray(arr).tobytes()).hexdigest()


def k2_tgsw_inputs():
    rng = G.rs(205)
    B = 2
    accum = G.torus32(rng, (B, K + 1, N))
    tr_sample = G.ff_numbers(rng, (B, K + 1, 2, N))
    bk = G.ff_numbers(rng, (2, K + 1, 2, K + 1, N))
    return accum, tr_sample, bk


def golden_tgsw_k2(out):
    from nufhe.tgsw_cpu import (
        tgsw_polynomial_decomp_trf_reference, tlwe_transformed_add_mul_to_trf_reference)
    accum, tr_sample, bk = k2_tgsw_inputs()
    B = accum.shape[0]
    dec = numpy.empty((B, K + 1, 2, N), numpy.int32)
    tgsw_polynomial_decomp_trf_reference(tgsw_params, (B,))(dec, accum)
    mac = numpy.empty((B, K + 1, N), numpy.uint64)
    tlwe_transformed_add_mul_to_trf_reference(tgsw_params, (B,), bk.shape[0], None)(mac, tr_sample, bk, 1)
    ext = accum.copy()
    TGswTransformedExternalMulReference(tgsw_params, (B,), bk.shape[0], None)(ext, bk, 0)
    out.update(decomp=dec, mac=mac, ext=ext)


def reference_keygen(seed):
    rng = numpy.random.RandomState(seed)
    n = 500
    lwe_key = rng.randint(0, 2, size=(n,), dtype=numpy.int32)
    tlwe_key = rng.randint(0, 2, size=(K, N), dtype=numpy.int32)
    bk_shape = (n, K + 1, 2)
    noises1 = rng.randint(-2**31, 2**31, size=bk_shape + (K, N), dtype=numpy.int32)
    noises2 = double_to_t32(rng.normal(size=bk_shape + (N,), scale=tlwe_params.min_noise))
    bk = numpy.empty(bk_shape + (K + 1, N), numpy.int32)
    cv = numpy.empty(bk_shape, numpy.float32)
    t = time.time()
    with numpy.errstate(over='ignore'):
        TLweEncryptZeroReference(tlwe_params, bk_shape, tlwe_params.min_noise, None)(
            bk, cv, tlwe_key, noises1, noises2)
        TGswAddMessageReference(tgsw_params, (n,))(bk, lwe_key)
    print('reference BK encrypt: %.1f s' % (time.time() - t), flush=True); t = time.time()
    bk_tr = prepare_for_mul_cpu(ntt_transform_ref(bk, i32_conversion=True))
    print('reference BK transform: %.1f s' % (time.time() - t), flush=True)
    ks_noise = params.in_out_params.min_noise
    noises_b = rng.normal(size=(K * N, 8, 3), scale=ks_noise)
    noises_b -= noises_b.mean()
    noises_b = double_to_t32(noises_b)
    noises_a = rng.randint(-2**31, 2**31, size=(K * N, 8, 3, n), dtype=numpy.int32)
    ks_a = numpy.empty((K * N, 8, 4, n), numpy.int32)
    ks_b = numpy.empty((K * N, 8, 4), numpy.int32)
    ks_cv = numpy.empty((K * N, 8, 4), numpy.float32)
    with numpy.errstate(over='ignore'):
        MakeLweKeyswitchKeyReference(K * N, n, 8, 2, ks_noise)(
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


def reference_bootstrap(x_a, x_b, bk_tr, ks, mu):
    """bootstrap(), nufhe/bootstrap.py:206-229 + :154-196 + :96-142 (the loop path) from the closures, k = 2."""
    B = x_b.shape[0]
    n = x_a.shape[-1]
    barb = numpy.empty((B,), numpy.int32)
    bara = numpy.empty((B, n), numpy.int32)
    Torus32ToPhaseReference((B,), 2 * N)(barb, x_b)
    Torus32ToPhaseReference((B, n), 2 * N)(bara, x_a)
    testvect = numpy.full((B, N), mu, numpy.int32)
    testvectbis = numpy.empty((B, N), numpy.int32)
    ShiftTorusPolynomialReference(N, (B,), (B,), invert_powers=True)(testvectbis, testvect, barb, 0)
    acc = numpy.empty((B, K + 1, N), numpy.int32)
    cv = numpy.empty((B,), numpy.float32)
    TLweNoiselessTrivialReference(tlwe_params, (B,))(acc, cv, testvectbis)
    shift = ShiftTorusPolynomialReference(N, (B, K + 1), (B, n), powers_view=True, minus_one=True)
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
    ea = numpy.empty((B, K * N), numpy.int32)
    eb = numpy.empty((B,), numpy.int32)
    TLweExtractLweSamplesReference(tlwe_params, (B,))(ea, eb, acc)
    ra = numpy.empty((B, n), numpy.int32)
    rb = numpy.empty((B,), numpy.int32)
    rcv = numpy.empty((B,), numpy.float32)
    with numpy.errstate(over='ignore'):
        LweKeyswitchReference(None, K * N, n, 8, 2)(ra, rb, rcv, ks[0], ks[1], ks[2], ea, eb)
    return (ra, rb), (ea, eb), acc


def golden_gate_k2(out):
    rng, lwe_key, tlwe_key, bk, bk_tr, ks = reference_keygen(G.GATE_SEED)
    c1 = reference_encrypt(rng, lwe_key, G.GATE_BITS_A[:2])
    c2 = reference_encrypt(rng, lwe_key, G.GATE_BITS_B[:2])
    with numpy.errstate(over='ignore'):
        t_a = (-c1[0] - c2[0]).astype(numpy.int32)
        t_b = (phase_to_t32(1, 8) - c1[1] - c2[1]).astype(numpy.int32)
    (nand_a, nand_b), (ext_a, ext_b), acc = reference_bootstrap(t_a, t_b, bk_tr, ks, phase_to_t32(1, 8))
    dec = numpy.empty((2,), numpy.int32)
    LweDecryptReference((2,), 500)(dec, nand_a, nand_b, lwe_key)
    print('k=2 NAND decrypts to', dec > 0, 'expected', ~(numpy.asarray(G.GATE_BITS_A[:2]) & numpy.asarray(G.GATE_BITS_B[:2])))
    out.update(
        seed=G.GATE_SEED, lwe_key_sha=sha(lwe_key), tlwe_key_sha=sha(tlwe_key), bk_raw_sha=sha(bk), bk_sha=sha(bk_tr),
        ks_a_sha=sha(ks[0]), ks_b_sha=sha(ks[1]), bk_row0=bk_tr[0], bk_row499=bk_tr[499],
        c1_a=c1[0], c1_b=c1[1], c2_a=c2[0], c2_b=c2[1],
        nand_a=nand_a, nand_b=nand_b, nand_ext_a=ext_a, nand_ext_b=ext_b, nand_acc=acc, nand_bits=(dec > 0))


if __name__ == '__main__':
    out = {}
    golden_tgsw_k2(out)
    numpy.savez_compressed(os.path.join(HERE, 'k2_small.npz'), **out)
    print('small k=2 vectors written', flush=True)
    golden_gate_k2(out)
    numpy.savez_compressed(os.path.join(HERE, 'k2.npz'), **out)
    print('wrote k2.npz:', {k: getattr(v, 'shape', None) for k, v in out.items()})
