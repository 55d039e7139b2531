"""Reference-closure goldens for BASELINE.md section 4: `gate_nand` on 32 ciphertexts and `gate_mux` on 4
ciphertext triples, every step through the UNMODIFIED reference's NumPy closures (the same composition as
make_golden.py::golden_gate, nufhe/gates.py:108-121 and :629-664, nufhe/bootstrap.py:206-229).

    python tests/golden/make_golden_gates32.py [processes]

The keys are generated once with the seed of gate.npz (so its key digests apply), then the 32 + 2 x 4 bootstraps are
split over `processes` forked workers (default: all cores); ~45 s of one core per bootstrap.  Output:
tests/golden/gates32.npz.  Runs in the build container only (needs /root/reference).
"""
import multiprocessing
import os
import sys
import time

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as M          # noqa: E402  (imports the reference through ref_bridge)
import gen_inputs as G           # noqa: E402

NAND_BATCH = 32
MUX_BATCH = 4
BITS_SEED = G.GATE_SEED + 32

_state = {}


def _bootstrap_job(job):
    kind, lo, hi = job
    bk_tr, ks = _state['bk_tr'], _state['ks']
    t_a, t_b = _state[kind]
    mu = M.phase_to_t32(1, 8)
    if kind == 'nand':
        (ra, rb), (ea, eb) = M.reference_bootstrap(t_a[lo:hi], t_b[lo:hi], bk_tr, ks, mu)
        return kind, lo, hi, ra, rb, ea, eb
    ea, eb = M.reference_bootstrap(t_a[lo:hi], t_b[lo:hi], bk_tr, ks, mu, no_keyswitch=True)
    return kind, lo, hi, None, None, ea, eb


def main():
    procs = int(sys.argv[1]) if len(sys.argv) > 1 else os.cpu_count()
    t0 = time.time()
    rng, lwe_key, tlwe_key, bk, bk_tr, ks = M.reference_keygen(G.GATE_SEED)
    # same draw order as gate.npz: three 4-ciphertext operands first, then the new operands
    for bits in (G.GATE_BITS_A, G.GATE_BITS_B, G.GATE_BITS_C):
        M.reference_encrypt(rng, lwe_key, bits)
    brng = numpy.random.RandomState(BITS_SEED)
    bits_a = brng.randint(0, 2, NAND_BATCH).astype(bool)
    bits_b = brng.randint(0, 2, NAND_BATCH).astype(bool)
    mux_bits = [numpy.array(x, bool) for x in ([1, 1, 0, 0], [1, 0, 1, 0], [0, 1, 1, 0])]
    d1 = M.reference_encrypt(rng, lwe_key, bits_a)
    d2 = M.reference_encrypt(rng, lwe_key, bits_b)
    m = [M.reference_encrypt(rng, lwe_key, b) for b in mux_bits]
    print('keys + inputs: %.0f s' % (time.time() - t0), flush=True)

    one8 = M.phase_to_t32(1, 8)
    and_const = M.phase_to_t32(-1, 8)
    with numpy.errstate(over='ignore'):
        # NAND: (0, 1/8) - a - b                                   gates.py:108-115
        _state['nand'] = ((-d1[0] - d2[0]).astype(numpy.int32), (one8 - d1[1] - d2[1]).astype(numpy.int32))
        # MUX: u1 = bootstrap((0,-1/8) + a + b), u2 = bootstrap((0,-1/8) - a + c), no key switch   gates.py:638-655
        _state['mux1'] = ((m[0][0] + m[1][0]).astype(numpy.int32), (and_const + m[0][1] + m[1][1]).astype(numpy.int32))
        _state['mux2'] = ((-m[0][0] + m[2][0]).astype(numpy.int32), (and_const - m[0][1] + m[2][1]).astype(numpy.int32))
    _state['bk_tr'], _state['ks'] = bk_tr, ks

    per = max(1, (NAND_BATCH + 2 * MUX_BATCH + procs - 1) // procs)
    jobs = [('nand', lo, min(lo + per, NAND_BATCH)) for lo in range(0, NAND_BATCH, per)]
    jobs += [('mux1', lo, min(lo + per, MUX_BATCH)) for lo in range(0, MUX_BATCH, per)]
    jobs += [('mux2', lo, min(lo + per, MUX_BATCH)) for lo in range(0, MUX_BATCH, per)]
    out = {'nand': [numpy.empty((NAND_BATCH, 500), numpy.int32), numpy.empty(NAND_BATCH, numpy.int32),
                    numpy.empty((NAND_BATCH, 1024), numpy.int32), numpy.empty(NAND_BATCH, numpy.int32)],
           'mux1': [None, None, numpy.empty((MUX_BATCH, 1024), numpy.int32), numpy.empty(MUX_BATCH, numpy.int32)],
           'mux2': [None, None, numpy.empty((MUX_BATCH, 1024), numpy.int32), numpy.empty(MUX_BATCH, numpy.int32)]}
    with multiprocessing.get_context('fork').Pool(procs) as pool:
        for kind, lo, hi, ra, rb, ea, eb in pool.imap_unordered(_bootstrap_job, jobs):
            o = out[kind]
            if ra is not None:
                o[0][lo:hi], o[1][lo:hi] = ra, rb
            o[2][lo:hi], o[3][lo:hi] = ea, eb
            print('done', kind, lo, hi, '%.0f s' % (time.time() - t0), flush=True)

    # MUX tail: (0, 1/8) + u1 + u2, then the key switch                       gates.py:657-664
    with numpy.errstate(over='ignore'):
        s_a = (out['mux1'][2] + out['mux2'][2]).astype(numpy.int32)
        s_b = (one8 + out['mux1'][3] + out['mux2'][3]).astype(numpy.int32)
        mux_a = numpy.empty((MUX_BATCH, 500), numpy.int32)
        mux_b = numpy.empty((MUX_BATCH,), numpy.int32)
        mux_cv = numpy.empty((MUX_BATCH,), numpy.float32)
        M.LweKeyswitchReference(None, 1024, 500, 8, 2)(mux_a, mux_b, mux_cv, ks[0], ks[1], ks[2], s_a, s_b)
    dec = numpy.empty((NAND_BATCH,), numpy.int32)
    M.LweDecryptReference((NAND_BATCH,), 500)(dec, out['nand'][0], out['nand'][1], lwe_key)
    assert ((dec > 0) == ~(bits_a & bits_b)).all(), 'reference NAND does not decrypt to the truth table'
    decm = numpy.empty((MUX_BATCH,), numpy.int32)
    M.LweDecryptReference((MUX_BATCH,), 500)(decm, mux_a, mux_b, lwe_key)
    want = numpy.where(mux_bits[0], mux_bits[1], mux_bits[2])
    assert ((decm > 0) == want).all(), 'reference MUX does not decrypt to the truth table'
    M.save('gates32', seed=G.GATE_SEED, bits_seed=BITS_SEED,
           lwe_key_sha=M.sha(lwe_key), bk_sha=M.sha(bk_tr), ks_a_sha=M.sha(ks[0]),
           bits_a=bits_a, bits_b=bits_b, in1_a=d1[0], in1_b=d1[1], in2_a=d2[0], in2_b=d2[1],
           nand_a=out['nand'][0], nand_b=out['nand'][1], nand_ext_a=out['nand'][2], nand_ext_b=out['nand'][3],
           mux_bits=numpy.stack(mux_bits),
           mux_in_a=numpy.stack([x[0] for x in m]), mux_in_b=numpy.stack([x[1] for x in m]),
           mux_u1_a=out['mux1'][2], mux_u1_b=out['mux1'][3], mux_u2_a=out['mux2'][2], mux_u2_b=out['mux2'][3],
           mux_a=mux_a, mux_b=mux_b)
    print('total %.0f s' % (time.time() - t0))


if __name__ == '__main__':
    main()
