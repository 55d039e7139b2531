"""The CPU oracle (oracle/nufhe_oracle.c) against the committed golden vectors, which were produced by
the reference's own NumPy closures (tests/golden/make_golden.py).  CPU only."""
import hashlib

import numpy
import pytest

import gen_inputs as G
from oracle import oracle as O


def sha(arr):
    return hashlib.sha256(numpy.ascontiguousarray(arr).tobytes()).hexdigest()


def test_reduction_selftest():
    assert O.selftest(seed=12345, n=200000) == 0


def test_arithmetic(golden):
    g = golden('arithmetic')
    a, b, s = G.arithmetic_inputs()
    assert (O.ff_mul(a, b) == g['mul']).all()
    assert (O.ff_add(a, b) == g['add']).all()
    assert (O.ff_sub(a, b) == g['sub']).all()
    assert (O.ff_mul_prepared(a, b) == g['mul_prepared']).all()
    assert (O.ff_prepare_for_mul(a) == g['prepare_for_mul']).all()
    assert (O.ff_lsh(a, s) == g['lsh']).all()


def test_ntt(golden):
    g = golden('ntt')
    x_i32, x_u64 = G.ntt_inputs()
    assert (O.ntt_forward_i32(x_i32) == g['fwd_i32']).all()
    assert (O.ntt_forward_u64(x_u64) == g['fwd_u64']).all()
    assert (O.ntt_inverse_u64(x_u64) == g['inv_u64']).all()
    assert (O.ntt_inverse_i32(x_u64) == g['inv_i32']).all()


def test_ntt_roundtrip_and_convolution():
    # same property as test/test_transform/test_computation.py:71-124 of the reference
    rng = G.rs(7)
    a = G.torus32(rng, (4, 1024))
    b = G.torus32(rng, (4, 1024), -1000, 1000)
    prod = O.ntt_inverse_i32(O.ff_mul(O.ntt_forward_i32(a), O.ntt_forward_i32(b)))
    a64, b64 = a.astype(object), b.astype(object)
    for q in range(4):
        full = numpy.convolve(a64[q], b64[q])
        neg = full[:1024].copy()
        neg[:1023] -= full[1024:]
        want = numpy.array([int(v) % 2**32 for v in neg], numpy.uint64).astype(numpy.uint32)
        assert (prod[q].view(numpy.uint32) == want).all()
    assert (O.ntt_inverse_i32(O.ntt_forward_i32(a)) == a).all()


def test_small_kernels(golden):
    g = golden('small')
    assert (O.t32_to_phase(G.modswitch_inputs(), 2048) == g['phase']).all()
    src, powers, bara = G.shift_inputs()
    assert (O.shift_torus_polynomial(src, powers) == g['shift_plain']).all()
    assert (O.shift_torus_polynomial(src, powers, invert_powers=True) == g['shift_inverted']).all()
    assert (O.shift_torus_polynomial(src, bara, 3, minus_one=True) == g['shift_minus_one']).all()
    ea, eb = O.tlwe_extract_lwe_samples(G.extract_inputs())
    assert (ea == g['extract_a']).all() and (eb == g['extract_b']).all()
    assert (O.tlwe_noiseless_trivial(src[:, 0, :]) == g['trivial']).all()
    a, b = G.linear_inputs()
    for name in ('nand', 'xor', 'andny'):
        num, den, sa, sb = O.GATE_TABLE[name]
        t_a, t_b = O.lwe_affine2(a, b, O.phase_to_t32(num, den), sa, sb)
        assert (t_a == g['lin_%s_a' % name]).all() and (t_b == g['lin_%s_b' % name]).all()


def test_tgsw(golden):
    g = golden('tgsw')
    accum_small, accum_full, tr_sample, bk = G.tgsw_inputs()
    assert (O.tgsw_decompose(accum_full) == g['decomp']).all()
    assert (O.tgsw_mac(tr_sample, bk, 1) == g['mac']).all()
    assert (O.tgsw_external_mul(accum_small, bk, 2) == g['ext_small']).all()
    assert (O.tgsw_external_mul(accum_full, bk, 0) == g['ext_full']).all()


def test_keyswitch(golden):
    g = golden('keyswitch')
    ks_a, ks_b, ks_cv, src_a, src_b = G.keyswitch_inputs()
    ra, rb, rcv = O.lwe_keyswitch(ks_a, ks_b, ks_cv, src_a, src_b)
    assert (ra == g['res_a']).all() and (rb == g['res_b']).all()
    assert numpy.allclose(rcv, g['res_cv'], rtol=1e-4, atol=1e-4)


@pytest.fixture(scope='module')
def keys():
    return O.OracleKeys(G.GATE_SEED)


def test_keygen_matches_reference(golden, keys):
    g = golden('gate')
    assert sha(keys.lwe_key) == str(g['lwe_key_sha'])
    assert sha(keys.tlwe_key) == str(g['tlwe_key_sha'])
    assert sha(keys.bk_raw) == str(g['bk_raw_sha'])
    assert sha(keys.bk) == str(g['bk_sha'])
    assert (keys.bk[0] == g['bk_row0']).all() and (keys.bk[499] == g['bk_row499']).all()
    assert sha(keys.ks_a) == str(g['ks_a_sha'])
    assert sha(keys.ks_b) == str(g['ks_b_sha'])


def test_gate_nand_matches_reference(golden, keys):
    g = golden('gate')
    c1 = keys.encrypt(G.GATE_BITS_A)
    c2 = keys.encrypt(G.GATE_BITS_B)
    c3 = keys.encrypt(G.GATE_BITS_C)
    assert (c1[0] == g['c1_a']).all() and (c1[1] == g['c1_b']).all()
    assert (c3[0] == g['c3_a']).all() and (c3[1] == g['c3_b']).all()
    sl = slice(0, 2)
    num, den, sa, sb = O.GATE_TABLE['nand']
    t = O.lwe_affine2((c1[0][sl], c1[1][sl]), (c2[0][sl], c2[1][sl]), O.phase_to_t32(num, den), sa, sb)
    ext = O.bootstrap(t[0], t[1], keys.bk, None)
    assert (ext[0] == g['nand_ext_a']).all() and (ext[1] == g['nand_ext_b']).all()
    out = O.gate_binary('nand', (c1[0][sl], c1[1][sl]), (c2[0][sl], c2[1][sl]), keys.bk, keys.ks)
    assert (out[0] == g['nand_a']).all() and (out[1] == g['nand_b']).all()
    assert (keys.decrypt(out) == g['nand_bits']).all()


def test_nand32_and_mux_match_reference(golden, keys):
    """BASELINE.md section 4's tier-0 set: 32 NAND ciphertexts and 4 MUX triples computed by the reference's own
    closures (tests/golden/gates32.npz, make_golden_gates32.py) -- the oracle reproduces the inputs (same RNG draw
    order after the three operands of gate.npz), the extracted samples and the final ciphertexts."""
    g = golden('gates32')
    assert sha(keys.lwe_key) == str(g['lwe_key_sha']) and sha(keys.bk) == str(g['bk_sha'])
    fresh = O.OracleKeys(G.GATE_SEED)
    for bits in (G.GATE_BITS_A, G.GATE_BITS_B, G.GATE_BITS_C):
        fresh.encrypt(bits)
    d1, d2 = fresh.encrypt(g['bits_a']), fresh.encrypt(g['bits_b'])
    m = [fresh.encrypt(g['mux_bits'][i]) for i in range(3)]
    assert (d1[0] == g['in1_a']).all() and (d1[1] == g['in1_b']).all() and (d2[0] == g['in2_a']).all()
    for i in range(3):
        assert (m[i][0] == g['mux_in_a'][i]).all() and (m[i][1] == g['mux_in_b'][i]).all()
    num, den, sa, sb = O.GATE_TABLE['nand']
    t = O.lwe_affine2(d1, d2, O.phase_to_t32(num, den), sa, sb)
    ext = O.bootstrap(t[0], t[1], keys.bk, None)
    assert (ext[0] == g['nand_ext_a']).all() and (ext[1] == g['nand_ext_b']).all()
    out = O.gate_binary('nand', d1, d2, keys.bk, keys.ks)
    assert (out[0] == g['nand_a']).all() and (out[1] == g['nand_b']).all()
    t1 = O.lwe_affine2(m[0], m[1], O.phase_to_t32(-1, 8), 1, 1)
    u1 = O.bootstrap(t1[0], t1[1], keys.bk, None)
    assert (u1[0] == g['mux_u1_a']).all() and (u1[1] == g['mux_u1_b']).all()
    mux = O.gate_mux(m[0], m[1], m[2], keys.bk, keys.ks)
    assert (mux[0] == g['mux_a']).all() and (mux[1] == g['mux_b']).all()


def test_all_gates_truth_tables(keys):
    a_bits = numpy.array(G.GATE_BITS_A)
    b_bits = numpy.array(G.GATE_BITS_B)
    c_bits = numpy.array(G.GATE_BITS_C)
    a, b, c = keys.encrypt(a_bits), keys.encrypt(b_bits), keys.encrypt(c_bits)
    truth = dict(
        nand=~(a_bits & b_bits), xor=a_bits ^ b_bits, xnor=~(a_bits ^ b_bits), nor=~(a_bits | b_bits),
        andny=~a_bits & b_bits, andyn=a_bits & ~b_bits, orny=~a_bits | b_bits, oryn=a_bits | ~b_bits)
    truth['or'] = a_bits | b_bits
    truth['and'] = a_bits & b_bits
    for name, want in truth.items():
        assert (keys.decrypt(O.gate_binary(name, a, b, keys.bk, keys.ks)) == want).all(), name
    assert (keys.decrypt(O.gate_mux(a, b, c, keys.bk, keys.ks)) == numpy.where(a_bits, b_bits, c_bits)).all()


def test_mask_size_2_steps(golden):
    """The general-k restatement of decomposition / MAC / external product against the reference's closures run with
    `NuFHEParameters(tlwe_mask_size=2)` (tests/golden/make_golden_k2.py), and against the k = 1 goldens."""
    g = golden('k2_small')
    rng = G.rs(205)
    accum = G.torus32(rng, (2, 3, 1024))
    tr = G.ff_numbers(rng, (2, 3, 2, 1024))
    bk = G.ff_numbers(rng, (2, 3, 2, 3, 1024))
    assert (O.tgsw_decompose_k(accum) == g['decomp']).all()
    assert (O.tgsw_mac_k(tr, bk[1]) == g['mac']).all()
    assert (O.tgsw_external_mul_k(accum, bk[0]) == g['ext']).all()
    g1 = golden('tgsw')
    accum_small, accum_full, tr_sample, bk1 = G.tgsw_inputs()
    assert (O.tgsw_decompose_k(accum_full) == g1['decomp']).all()
    assert (O.tgsw_mac_k(tr_sample, bk1[1]) == g1['mac']).all()
    assert (O.tgsw_external_mul_k(accum_full, bk1[0]) == g1['ext_full']).all()


def test_mask_size_2_keys_and_gate(golden):
    """Oracle key generation and the whole bootstrap for tlwe_mask_size = 2 against the reference's closures
    (tests/golden/k2.npz): key digests, ciphertexts, the final accumulator, the extracted and the key-switched sample."""
    import hashlib
    g = golden('k2')
    keys = O.OracleKeys(int(g['seed']), mask_size=2)

    def sha(a):
        return hashlib.sha256(numpy.ascontiguousarray(a).tobytes()).hexdigest()
    assert sha(keys.lwe_key) == str(g['lwe_key_sha']) and sha(keys.tlwe_key) == str(g['tlwe_key_sha'])
    assert sha(keys.bk_raw) == str(g['bk_raw_sha']) and sha(keys.bk) == str(g['bk_sha'])
    assert sha(keys.ks_a) == str(g['ks_a_sha']) and sha(keys.ks_b) == str(g['ks_b_sha'])
    c1, c2 = keys.encrypt(G.GATE_BITS_A[:2]), keys.encrypt(G.GATE_BITS_B[:2])
    assert (c1[0] == g['c1_a']).all() and (c1[1] == g['c1_b']).all() and (c2[0] == g['c2_a']).all()
    t_a, t_b = O.lwe_affine2(c1, c2, O.phase_to_t32(1, 8), -1, -1)
    (ra, rb), acc = O.bootstrap_k(t_a, t_b, keys.bk, keys.ks)
    assert (acc == g['nand_acc']).all()
    assert (ra == g['nand_a']).all() and (rb == g['nand_b']).all()
    (ea, eb), _ = O.bootstrap_k(t_a, t_b, keys.bk, None)
    assert (ea == g['nand_ext_a']).all() and (eb == g['nand_ext_b']).all()
    assert (keys.decrypt((ra, rb)) == g['nand_bits']).all()
    # and the same composition at k = 1 equals the fused restatement
    k1 = O.OracleKeys(G.GATE_SEED)
    a, b = k1.encrypt([True, False]), k1.encrypt([True, True])
    t_a, t_b = O.lwe_affine2(a, b, O.phase_to_t32(1, 8), -1, -1)
    (ra, rb), _ = O.bootstrap_k(t_a, t_b, k1.bk, k1.ks)
    want = O.bootstrap(t_a, t_b, k1.bk, k1.ks)
    assert (ra == want[0]).all() and (rb == want[1]).all()
