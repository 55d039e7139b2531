"""Gate-level parity on the GPU: fused bootstrap + key switch through the C ABI against
(1) the committed output of the reference's own closures (tests/golden/gate.npz) and
(2) the CPU oracle on larger batches, plus decryption truth tables (test/test_gates.py:40-85)."""
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


@pytest.fixture(scope='module')
def keys():
    return O.OracleKeys(G.GATE_SEED)


@pytest.fixture(scope='module')
def dev_keys(eng, keys):
    bk_int = eng.bk_prepare(eng.to_device(keys.bk))
    ks = (eng.to_device(keys.ks_a), eng.to_device(keys.ks_b), eng.to_device(keys.ks_cv))
    return bk_int, ks


def gpu_gate(eng, dev_keys, name, a, b):
    bk_int, ks = dev_keys
    num, den, sa, sb = O.GATE_TABLE[name]
    da = (eng.to_device(a[0]), eng.to_device(a[1]))
    db = (eng.to_device(b[0]), eng.to_device(b[1]))
    ext = eng.bootstrap_extract(da, db, O.phase_to_t32(num, den), sa, sb, O.MU, bk_int)
    ra, rb, _ = eng.keyswitch(ks, ext)
    return (eng.to_host(ext[0]), eng.to_host(ext[1])), (eng.to_host(ra), eng.to_host(rb))


def test_nand_matches_reference_closures(eng, keys, dev_keys, golden):
    g = golden('gate')
    c1 = (g['c1_a'][:2], g['c1_b'][:2])
    c2 = (g['c2_a'][:2], g['c2_b'][:2])
    ext, out = gpu_gate(eng, dev_keys, 'nand', c1, c2)
    assert (ext[0] == g['nand_ext_a']).all() and (ext[1] == g['nand_ext_b']).all()
    assert (out[0] == g['nand_a']).all() and (out[1] == g['nand_b']).all()
    assert (keys.decrypt(out) == g['nand_bits']).all()


@pytest.mark.parametrize('name,batch', [('nand', 1), ('nand', 37), ('xor', 8), ('andny', 5)])
def test_binary_gates_vs_oracle(eng, keys, dev_keys, name, batch):
    rng = G.rs(400 + batch)
    bits_a, bits_b = rng.randint(0, 2, batch).astype(bool), rng.randint(0, 2, batch).astype(bool)
    a, b = keys.encrypt(bits_a), keys.encrypt(bits_b)
    ext, out = gpu_gate(eng, dev_keys, name, a, b)
    want = O.gate_binary(name, a, b, keys.bk, keys.ks)
    assert (out[0] == want[0]).all() and (out[1] == want[1]).all()
    truth = dict(nand=~(bits_a & bits_b), xor=bits_a ^ bits_b, andny=~bits_a & bits_b)[name]
    assert (keys.decrypt(out) == truth).all()


def test_mux_vs_oracle(eng, keys, dev_keys):
    bk_int, ks = dev_keys
    rng = G.rs(450)
    B = 6
    bits = [rng.randint(0, 2, B).astype(bool) for _ in range(3)]
    a, b, c = (keys.encrypt(x) for x in bits)
    d = [(eng.to_device(x[0]), eng.to_device(x[1])) for x in (a, b, c)]
    and_const = O.phase_to_t32(-1, 8)
    u1 = eng.bootstrap_extract(d[0], d[1], and_const, 1, 1, O.MU, bk_int)
    u2 = eng.bootstrap_extract(d[0], d[2], and_const, -1, 1, O.MU, bk_int)
    ra, rb, _ = eng.keyswitch(ks, u1, u2, c=O.phase_to_t32(1, 8))
    want = O.gate_mux(a, b, c, keys.bk, keys.ks)
    assert (eng.to_host(ra) == want[0]).all() and (eng.to_host(rb) == want[1]).all()
    assert (keys.decrypt((eng.to_host(ra), eng.to_host(rb))) == numpy.where(bits[0], bits[1], bits[2])).all()


def test_nand_more_than_one_wave_vs_oracle(eng, keys, dev_keys):
    """600 ciphertexts = 300 CTAs > the 296 co-resident ones: exercises CTA turnover; every (a, b) is compared."""
    rng = G.rs(470)
    B = 600
    bits_a, bits_b = rng.randint(0, 2, B).astype(bool), rng.randint(0, 2, B).astype(bool)
    a, b = keys.encrypt(bits_a), keys.encrypt(bits_b)
    ext, out = gpu_gate(eng, dev_keys, 'nand', a, b)
    want = O.gate_binary('nand', a, b, keys.bk, keys.ks)
    assert (out[0] == want[0]).all() and (out[1] == want[1]).all()
    assert (keys.decrypt(out) == ~(bits_a & bits_b)).all()


@pytest.mark.parametrize('batch', [4096, 16384])
def test_full_size_batches_decrypt_to_truth_table(eng, keys, dev_keys, batch):
    """BASELINE.json batch sizes: size-independent property (encrypt -> gate -> decrypt == truth table) for
    NAND and MUX, plus determinism (two runs give identical ciphertexts) and batch-independence (the first
    37 outputs equal those of a 37-ciphertext run)."""
    bk_int, ks = dev_keys
    rng = G.rs(480 + batch)
    bits = [rng.randint(0, 2, batch).astype(bool) for _ in range(3)]
    cts = [keys.encrypt(x) for x in bits]
    d = [(eng.to_device(x[0]), eng.to_device(x[1])) for x in cts]
    num, den, sa, sb = O.GATE_TABLE['nand']
    ext = eng.bootstrap_extract(d[0], d[1], O.phase_to_t32(num, den), sa, sb, O.MU, bk_int)
    ra, rb, _ = eng.keyswitch(ks, ext)
    out = (eng.to_host(ra), eng.to_host(rb))
    assert (keys.decrypt(out) == ~(bits[0] & bits[1])).all()
    ext2 = eng.bootstrap_extract(d[0], d[1], O.phase_to_t32(num, den), sa, sb, O.MU, bk_int)
    ra2, rb2, _ = eng.keyswitch(ks, ext2)
    assert torch.equal(ra, ra2) and torch.equal(rb, rb2)
    small = [(x[0][:37].contiguous(), x[1][:37].contiguous()) for x in d[:2]]
    ext3 = eng.bootstrap_extract(small[0], small[1], O.phase_to_t32(num, den), sa, sb, O.MU, bk_int)
    ra3, rb3, _ = eng.keyswitch(ks, ext3)
    assert torch.equal(ra[:37], ra3) and torch.equal(rb[:37], rb3)
    and_const = O.phase_to_t32(-1, 8)
    u1, u2 = eng.bootstrap_extract2((d[0], d[1], and_const, 1, 1), (d[0], d[2], and_const, -1, 1), O.MU, bk_int)
    ma, mb, _ = eng.keyswitch(ks, u1, u2, c=O.phase_to_t32(1, 8))
    assert (keys.decrypt((eng.to_host(ma), eng.to_host(mb))) == numpy.where(bits[0], bits[1], bits[2])).all()


def test_both_cta_shapes_give_the_same_bits(keys, monkeypatch):
    """The fused kernel has two CTA shapes (2 ciphertexts per 256 threads; 1 ciphertext per 256 threads for batches
    that fit one wave).  Force each shape for the same inputs, including a ragged batch and one larger than a wave
    of wide CTAs; both must equal the oracle."""
    from nufhe_b200.engine import Engine
    rng = G.rs(480)
    outs = {}
    for B in (1, 5, 301):
        bits_a, bits_b = rng.randint(0, 2, B).astype(bool), rng.randint(0, 2, B).astype(bool)
        a, b = keys.encrypt(bits_a), keys.encrypt(bits_b)
        want = O.gate_binary('nand', a, b, keys.bk, keys.ks) if B <= 5 else None
        for wide_max in ('0', '1000000'):
            monkeypatch.setenv('NUFHE_B200_WIDE_MAX', wide_max)
            eng = Engine()
            dk = (eng.bk_prepare(eng.to_device(keys.bk)),
                  (eng.to_device(keys.ks_a), eng.to_device(keys.ks_b), eng.to_device(keys.ks_cv)))
            ext, out = gpu_gate(eng, dk, 'nand', a, b)
            outs[(B, wide_max)] = (ext, out)
            if want is not None:
                assert (out[0] == want[0]).all() and (out[1] == want[1]).all(), (B, wide_max)
            assert (keys.decrypt(out) == ~(bits_a & bits_b)).all()
        for x, y in zip(outs[(B, '0')], outs[(B, '1000000')]):
            assert (x[0] == y[0]).all() and (x[1] == y[1]).all(), B
