"""The Python host layer end to end on the CPU: the real nufhe_b200 package driven through a test double of the
engine (tests/fake_engine.py, backed by the oracle) and compared with the reference's golden vectors.

What this pins without a GPU: key generation in the reference's RNG order (k = 1 and k = 2), encryption, the fused
gate path's host side, the literal multi-kernel sequence (`single_kernel_bootstrap=False`: mod-switch, test vector,
500 x (rotate, external product, add), extraction, key switch), the k = 2 flow with its separate external-product
steps, broadcasting, gate_mux, serialization.  The CUDA kernels themselves are tested in test_gpu_*.py."""
import hashlib

import numpy
import pytest

import gen_inputs as G
from fake_engine import FakeEngine


def host(t, unsigned=False):
    a = t.cpu().numpy()
    return a.view(numpy.uint64) if unsigned else a


def sha(t, unsigned=False):
    return hashlib.sha256(numpy.ascontiguousarray(host(t, unsigned)).tobytes()).hexdigest()


@pytest.fixture(scope='module')
def nufhe():
    import nufhe_b200
    return nufhe_b200


@pytest.fixture(scope='module')
def k1(nufhe):
    ctx = nufhe.Context(rng=nufhe.DeterministicRNG(G.GATE_SEED), thread=FakeEngine())
    sk, ck = ctx.make_key_pair()
    return ctx, sk, ck


def test_keys_and_gates_k1_against_reference_golden(nufhe, k1, golden):
    ctx, sk, ck = k1
    g = golden('gate')
    assert sha(sk.lwe_key.key) == str(g['lwe_key_sha'])
    assert sha(ck.bootstrap_key.tgsw.samples.a.coeffs, True) == str(g['bk_sha'])
    ks = ck.keyswitch_key.lwe
    assert sha(ks.a) == str(g['ks_a_sha']) and sha(ks.b) == str(g['ks_b_sha'])
    c1, c2, c3 = (ctx.encrypt(sk, b) for b in (G.GATE_BITS_A, G.GATE_BITS_B, G.GATE_BITS_C))
    for c, name in ((c1, 'c1'), (c2, 'c2'), (c3, 'c3')):
        assert (host(c.a) == g[name + '_a']).all() and (host(c.b) == g[name + '_b']).all()
    vm_fused = ctx.make_virtual_machine(ck)
    vm_steps = ctx.make_virtual_machine(ck, perf_params=nufhe.PerformanceParameters(ck.params, single_kernel_bootstrap=False))
    thr = ctx.thread
    adds_before = thr.calls.get('tlwe_add_to', 0)                    # key generation sums the TLWE bodies with it too
    for vm, counter in ((vm_fused, 'bootstrap_extract'), (vm_steps, 'shift_torus_polynomial')):
        before = thr.calls.get(counter, 0)
        r = vm.gate_nand(c1[:2], c2[:2])
        assert thr.calls.get(counter, 0) > before                    # the intended path ran
        assert (host(r.a) == g['nand_a']).all() and (host(r.b) == g['nand_b']).all()
        assert (ctx.decrypt(sk, r) == g['nand_bits']).all()
    # 500 steps x 3 launches + the test-vector rotation
    assert thr.calls['shift_torus_polynomial'] == 501 and thr.calls['tlwe_add_to'] - adds_before == 500
    # MUX on both paths gives the same ciphertext and the right bits
    m1, m2 = vm_fused.gate_mux(c1, c2, c3), vm_steps.gate_mux(c1, c2, c3)
    assert (host(m1.a) == host(m2.a)).all() and (host(m1.b) == host(m2.b)).all()
    a, b, c = (numpy.array(x) for x in (G.GATE_BITS_A, G.GATE_BITS_B, G.GATE_BITS_C))
    assert (ctx.decrypt(sk, m1) == numpy.where(a, b, c)).all()


def test_truth_tables_broadcasting_and_serialization(nufhe, k1):
    ctx, sk, ck = k1
    vm = ctx.make_virtual_machine(ck)
    a = numpy.array([[True], [False]])                 # (2, 1) against (2,) -> (2, 2)
    b = numpy.array([True, False])
    ca, cb = ctx.encrypt(sk, a), ctx.encrypt(sk, b)
    for name, fn in (('gate_and', lambda x, y: x & y), ('gate_or', lambda x, y: x | y), ('gate_xor', lambda x, y: x ^ y),
                     ('gate_andny', lambda x, y: ~x & y), ('gate_oryn', lambda x, y: x | ~y)):
        r = getattr(vm, name)(ca, cb)
        assert tuple(r.shape) == (2, 2)
        assert (ctx.decrypt(sk, r) == fn(a, b)).all(), name
    assert (ctx.decrypt(sk, vm.gate_not(ca)) == ~a).all()
    with pytest.raises(ValueError):
        vm.gate_and(ctx.encrypt(sk, numpy.ones(3, bool)), cb)
    ck2 = ctx.load_cloud_key(ck.dumps())
    assert ck2 == ck
    ct2 = ctx.load_ciphertext(ca.dumps())
    assert ct2 == ca


def test_mask_size_2_flow_against_reference_golden(nufhe, golden):
    """`tlwe_mask_size=2` through the real host code: the keys (digests), the ciphertexts and the bits of one complete
    gate_nand must equal what the reference's closures produced (tests/golden/k2.npz)."""
    g = golden('k2')
    ctx = nufhe.Context(rng=nufhe.DeterministicRNG(int(g['seed'])), thread=FakeEngine())
    sk, ck = ctx.make_key_pair(tlwe_mask_size=2)
    bk = ck.bootstrap_key.tgsw.samples.a.coeffs
    assert tuple(bk.shape) == (500, 3, 2, 3, 1024)
    assert sha(sk.lwe_key.key) == str(g['lwe_key_sha']) and sha(bk, True) == str(g['bk_sha'])
    ks = ck.keyswitch_key.lwe
    assert sha(ks.a) == str(g['ks_a_sha']) and sha(ks.b) == str(g['ks_b_sha'])
    c1, c2 = ctx.encrypt(sk, G.GATE_BITS_A[:2]), ctx.encrypt(sk, G.GATE_BITS_B[:2])
    assert (host(c1.a) == g['c1_a']).all() and (host(c2.b) == g['c2_b']).all()
    vm = ctx.make_virtual_machine(ck)
    assert not vm.perf_params.single_kernel_bootstrap
    r = vm.gate_nand(c1, c2)
    assert ctx.thread.calls['tgsw_mac'] == 500 and ctx.thread.calls.get('external_product', 0) == 0
    assert (host(r.a) == g['nand_a']).all() and (host(r.b) == g['nand_b']).all()
    assert (ctx.decrypt(sk, r) == g['nand_bits']).all()
    with pytest.raises(ValueError):
        ctx.make_virtual_machine(ck, perf_params=nufhe.PerformanceParameters(ck.params, single_kernel_bootstrap=True))


def test_uint_min_views_roll_concatenate(nufhe, k1):
    """The callers around the path: the `uint_min` circuit on views, `roll`, `concatenate`,
    `__setitem__` -- host logic only, so it runs on the CPU double."""
    from nufhe_b200.operators_integer import uint_min, uintarray_to_bitarray, bitarray_to_uintarray
    ctx, sk, ck = k1
    vm = ctx.make_virtual_machine(ck)
    xs = numpy.array([17, 200, 3, 128], numpy.uint8)
    ys = numpy.array([17, 100, 5, 127], numpy.uint8)
    ca, cb = ctx.encrypt(sk, uintarray_to_bitarray(xs)), ctx.encrypt(sk, uintarray_to_bitarray(ys))
    answer = vm.empty_ciphertext((4, 8))
    uint_min(ctx.thread, ck, answer, ca, cb, perf_params=vm.perf_params)
    assert (bitarray_to_uintarray(ctx.decrypt(sk, answer)) == numpy.minimum(xs, ys)).all()
    bits = uintarray_to_bitarray(xs)
    rolled = ca.copy()
    rolled.roll(3, axis=-1)
    assert (ctx.decrypt(sk, rolled) == numpy.roll(bits, 3, axis=-1)).all()
    both = nufhe.concatenate([ca, cb], axis=0)
    assert tuple(both.shape) == (8, 8)
    assert (ctx.decrypt(sk, both) == numpy.concatenate([bits, uintarray_to_bitarray(ys)], axis=0)).all()
    both[0:4] = cb
    assert (ctx.decrypt(sk, both[:4]) == uintarray_to_bitarray(ys)).all()
    view = ca[1:3, ::2]
    assert tuple(view.shape) == (2, 4)
    r = vm.gate_not(view)
    assert (ctx.decrypt(sk, r) == ~bits[1:3, ::2]).all()


def test_transform_interface_on_the_double():
    """nufhe's `Transform` / `ForwardTransform` / `InverseTransform` wrappers (nufhe_b200/transform.py): shapes,
    compile(), round trip, error behaviour -- host logic, the kernels are tested in test_gpu_kernels.py."""
    import torch
    from nufhe_b200.transform import Transform, ForwardTransform, InverseTransform, get_transform, transformed_dtype
    eng = FakeEngine()
    x = G.torus32(G.rs(1), (2, 3, 1024))
    fwd = ForwardTransform((2, 3), 1024, None).compile(eng)
    inv = InverseTransform((2, 3), 1024, None).compile(eng)
    tr = eng.empty((2, 3, 1024), torch.int64)
    fwd(tr, eng.to_device(x))
    back = eng.empty((2, 3, 1024), torch.int32)
    inv(back, tr)
    assert (back.numpy() == x).all() and transformed_dtype() == numpy.uint64
    with pytest.raises(ValueError):
        Transform(None, (2,))(tr, tr)                   # not compiled
    with pytest.raises(ValueError):
        fwd(tr, eng.to_device(x[:1]))                   # wrong batch shape
    with pytest.raises(ValueError):
        get_transform('FFT')


def test_mask_size_2_gates_against_the_general_oracle(nufhe):
    """Beyond the golden: a small batch of k = 2 gates through the host layer equals the oracle's general-k bootstrap
    (oracle.bootstrap_k, itself pinned to the reference's k = 2 closures in test_oracle.py)."""
    from oracle import oracle as O
    seed = 424242
    ctx = nufhe.Context(rng=nufhe.DeterministicRNG(seed), thread=FakeEngine())
    sk, ck = ctx.make_key_pair(tlwe_mask_size=2)
    keys = O.OracleKeys(seed, mask_size=2)
    assert (host(ck.bootstrap_key.tgsw.samples.a.coeffs, True) == keys.bk).all()
    assert (host(ck.keyswitch_key.lwe.a) == keys.ks_a).all() and (host(ck.keyswitch_key.lwe.b) == keys.ks_b).all()
    a = numpy.array([True, False, True])
    b = numpy.array([False, False, True])
    ca, cb = ctx.encrypt(sk, a), ctx.encrypt(sk, b)
    oa, ob = keys.encrypt(a), keys.encrypt(b)
    assert (host(ca.a) == oa[0]).all() and (host(cb.b) == ob[1]).all()
    vm = ctx.make_virtual_machine(ck)
    for name, (num, den, sa, sb) in (('gate_or', O.GATE_TABLE['or']), ('gate_andyn', O.GATE_TABLE['andyn'])):
        r = getattr(vm, name)(ca, cb)
        t_a, t_b = O.lwe_affine2(oa, ob, O.phase_to_t32(num, den), sa, sb)
        (wa, wb), _ = O.bootstrap_k(t_a, t_b, keys.bk, keys.ks)
        assert (host(r.a) == wa).all() and (host(r.b) == wb).all(), name
    assert (ctx.decrypt(sk, vm.gate_or(ca, cb)) == (a | b)).all()
