"""The nufhe-compatible Python surface on the GPU: seeded keys equal the reference's (through the
oracle, itself pinned to the reference), gates decrypt to truth tables, broadcasting / views /
serialization / error behaviour follow test/test_api_high_level.py, test_api_low_level.py and
test_gates.py of the reference."""
import io

import numpy
import pytest
import torch

import gen_inputs as G
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def nufhe():
    import nufhe_b200
    return nufhe_b200


@pytest.fixture(scope='module')
def ctx(nufhe):
    return nufhe.Context(rng=nufhe.DeterministicRNG(G.GATE_SEED))


@pytest.fixture(scope='module')
def key_pair(ctx):
    return ctx.make_key_pair()


@pytest.fixture(scope='module')
def okeys():
    return O.OracleKeys(G.GATE_SEED)


def host(t, unsigned=False):
    a = t.cpu().numpy()
    return a.view(numpy.uint64) if unsigned else a


def test_seeded_keys_equal_reference_keys(key_pair, okeys):
    sk, ck = key_pair
    assert (host(sk.lwe_key.key) == okeys.lwe_key).all()
    assert (host(ck.bootstrap_key.tgsw.samples.a.coeffs, True) == okeys.bk).all()
    ks = ck.keyswitch_key.lwe
    assert (host(ks.a) == okeys.ks_a).all() and (host(ks.b) == okeys.ks_b).all()
    assert numpy.allclose(host(ks.current_variances), okeys.ks_cv)


def test_encrypt_gate_decrypt_bit_exact(ctx, key_pair, okeys, nufhe):
    sk, ck = key_pair
    vm = ctx.make_virtual_machine(ck)
    c1, c2 = ctx.encrypt(sk, G.GATE_BITS_A), ctx.encrypt(sk, G.GATE_BITS_B)
    o1, o2 = okeys.encrypt(G.GATE_BITS_A), okeys.encrypt(G.GATE_BITS_B)
    assert (host(c1.a) == o1[0]).all() and (host(c1.b) == o1[1]).all()
    assert (host(c2.a) == o2[0]).all() and (host(c2.b) == o2[1]).all()
    r = vm.gate_nand(c1, c2)
    want = O.gate_binary('nand', o1, o2, okeys.bk, okeys.ks)
    assert (host(r.a) == want[0]).all() and (host(r.b) == want[1]).all()
    assert (ctx.decrypt(sk, r) == ~(numpy.array(G.GATE_BITS_A) & numpy.array(G.GATE_BITS_B))).all()


ALL_BINARY = ['nand', 'or', 'and', 'xor', 'xnor', 'nor', 'andny', 'andyn', 'orny', 'oryn']


def truth(name, a, b):
    return dict(nand=~(a & b), xor=a ^ b, xnor=~(a ^ b), nor=~(a | b), andny=~a & b, andyn=a & ~b,
                orny=~a | b, oryn=a | ~b, **{'or': a | b, 'and': a & b})[name]


@pytest.mark.parametrize('name', ALL_BINARY)
def test_binary_gate_truth_tables(ctx, key_pair, name):
    sk, ck = key_pair
    vm = ctx.make_virtual_machine(ck)
    rng = numpy.random.RandomState(3)
    a, b = rng.randint(0, 2, 32).astype(bool), rng.randint(0, 2, 32).astype(bool)
    r = getattr(vm, 'gate_' + name)(ctx.encrypt(sk, a), ctx.encrypt(sk, b))
    assert (ctx.decrypt(sk, r) == truth(name, a, b)).all()


def test_mux_not_copy_constant(ctx, key_pair):
    sk, ck = key_pair
    vm = ctx.make_virtual_machine(ck)
    rng = numpy.random.RandomState(4)
    a, b, c = (rng.randint(0, 2, (2, 8)).astype(bool) for _ in range(3))
    ca, cb, cc = ctx.encrypt(sk, a), ctx.encrypt(sk, b), ctx.encrypt(sk, c)
    assert (ctx.decrypt(sk, vm.gate_mux(ca, cb, cc)) == numpy.where(a, b, c)).all()
    assert (ctx.decrypt(sk, vm.gate_not(ca)) == ~a).all()
    assert (ctx.decrypt(sk, vm.gate_copy(ca)) == a).all()
    dest = vm.empty_ciphertext((2, 8))
    vm.gate_constant(c[1].tolist(), dest=dest)
    assert (ctx.decrypt(sk, dest) == numpy.broadcast_to(c[1], (2, 8))).all()


def test_mux_bit_exact_vs_oracle(ctx, key_pair, okeys):
    """vm.gate_mux (two bootstraps in one launch + fused key switch) against the oracle's gate_mux."""
    sk, ck = key_pair
    vm = ctx.make_virtual_machine(ck)
    rng = numpy.random.RandomState(6)
    bits = [rng.randint(0, 2, 7).astype(bool) for _ in range(3)]
    cts = [ctx.encrypt(sk, b) for b in bits]
    r = vm.gate_mux(*cts)
    want = O.gate_mux(*[(host(c.a), host(c.b)) for c in cts], okeys.bk, okeys.ks)
    assert (host(r.a) == want[0]).all() and (host(r.b) == want[1]).all()
    assert (ctx.decrypt(sk, r) == numpy.where(bits[0], bits[1], bits[2])).all()


def test_broadcasting_views_and_dest(ctx, key_pair):
    sk, ck = key_pair
    vm = ctx.make_virtual_machine(ck)
    rng = numpy.random.RandomState(5)
    a, b = rng.randint(0, 2, (3, 1)).astype(bool), rng.randint(0, 2, (4,)).astype(bool)
    r = vm.gate_and(ctx.encrypt(sk, a), ctx.encrypt(sk, b))
    assert r.shape == (3, 4) and (ctx.decrypt(sk, r) == (a & b)).all()
    # strided views as inputs and as output (test_gates.py:514-559)
    x, y = rng.randint(0, 2, (4, 6)).astype(bool), rng.randint(0, 2, (4, 6)).astype(bool)
    cx, cy = ctx.encrypt(sk, x), ctx.encrypt(sk, y)
    dest = vm.empty_ciphertext((4, 6))
    vm.gate_constant(numpy.zeros((4, 6), bool), dest=dest)
    vm.gate_xor(cx[1:3, ::2], cy[1:3, ::2], dest=dest[1:3, ::2])
    got = ctx.decrypt(sk, dest)
    want = numpy.zeros((4, 6), bool)
    want[1:3, ::2] = x[1:3, ::2] ^ y[1:3, ::2]
    assert (got == want).all()
    # roll / concatenate / setitem / copy
    import nufhe_b200 as nufhe
    cat = nufhe.concatenate([cx, cy], axis=1)
    assert cat.shape == (4, 12) and (ctx.decrypt(sk, cat) == numpy.concatenate([x, y], axis=1)).all()
    cz = cx.copy()
    cz.roll(2, axis=-1)
    assert (ctx.decrypt(sk, cz) == numpy.roll(x, 2, axis=-1)).all()
    cz[0] = cy[3]
    assert (ctx.decrypt(sk, cz)[0] == y[3]).all()
    with pytest.raises(ValueError):
        cz[0] = 5


def test_errors(ctx, key_pair, nufhe):
    sk, ck = key_pair
    vm = ctx.make_virtual_machine(ck)
    c3, c4 = ctx.encrypt(sk, [1, 0, 1]), ctx.encrypt(sk, [1, 0, 1, 1])
    with pytest.raises(ValueError):
        vm.gate_nand(c3, c4)
    with pytest.raises(ValueError):
        vm.gate_nand(c3, c3, dest=vm.empty_ciphertext((2,)))
    with pytest.raises(AttributeError):
        vm.nand
    with pytest.raises(ValueError):
        nufhe.Context(api='OpenCL')


def test_serialization_roundtrip(ctx, key_pair, nufhe):
    sk, ck = key_pair
    sk2 = ctx.load_secret_key(sk.dumps())
    assert sk2 == sk
    ct = ctx.encrypt(sk, [1, 0, 0, 1])
    ct2 = ctx.load_ciphertext(ct.dumps())
    assert ct2 == ct
    f = io.BytesIO()
    ck.dump(f)
    f.seek(0)
    ck2 = ctx.load_cloud_key(f)
    assert ck2 == ck
    vm = ctx.make_virtual_machine(ck2)
    assert (ctx.decrypt(sk2, vm.gate_or(ct2, ct2)) == [True, False, False, True]).all()


def test_low_level_api(key_pair, okeys, nufhe):
    from nufhe_b200.engine import Engine
    thr = Engine(0)
    rng = nufhe.DeterministicRNG(G.GATE_SEED)
    sk, ck = nufhe.make_key_pair(thr, rng)
    assert (host(sk.lwe_key.key) == okeys.lwe_key).all()
    ct = nufhe.encrypt(thr, rng, sk, [True, False])
    res = nufhe.empty_ciphertext(thr, ck.params, (2,))
    nufhe.gate_nand(thr, ck, res, ct, ct)
    assert (nufhe.decrypt(thr, sk, res) == [False, True]).all()
    # bootstrap seam (bootstrap.py:206-209) incl. no_keyswitch, and the explicit blind-rotate seam
    from nufhe_b200.bootstrap import bootstrap
    from nufhe_b200.lwe import LweSampleArray
    ext = LweSampleArray.empty(thr, ck.params.tgsw_params.tlwe_params.extracted_lweparams, (2,))
    bootstrap(thr, ext, ck.bootstrap_key, ck.keyswitch_key, 2**29, ct, no_keyswitch=True)
    want = O.bootstrap(host(ct.a), host(ct.b), okeys.bk, None)
    assert (host(ext.a) == want[0]).all() and (host(ext.b) == want[1]).all()


def test_uint_min_circuit(ctx, key_pair):
    """operators_integer.uint_min (test_gates.py:248-249 of the reference): 8-bit encrypted minimum."""
    from nufhe_b200.operators_integer import uint_min, uintarray_to_bitarray, bitarray_to_uintarray
    sk, ck = key_pair
    vm = ctx.make_virtual_machine(ck)
    rng = numpy.random.RandomState(8)
    xs, ys = rng.randint(0, 256, 6).astype(numpy.uint8), rng.randint(0, 256, 6).astype(numpy.uint8)
    xs[0], ys[0] = 17, 17
    ca, cb = ctx.encrypt(sk, uintarray_to_bitarray(xs)), ctx.encrypt(sk, uintarray_to_bitarray(ys))
    answer = vm.empty_ciphertext((6, 8))
    uint_min(ctx.thread, ck, answer, ca, cb, perf_params=vm.perf_params)
    got = bitarray_to_uintarray(ctx.decrypt(sk, answer))
    assert (got == numpy.minimum(xs, ys)).all()


def test_empty_batch_is_a_no_op(ctx, key_pair, nufhe):
    """Zero ciphertexts: every gate returns an empty ciphertext without launching anything."""
    sk, ck = key_pair
    vm = ctx.make_virtual_machine(ck)
    a = ctx.encrypt(sk, numpy.zeros((0,), bool))
    b = ctx.encrypt(sk, numpy.zeros((0,), bool))
    assert vm.gate_nand(a, b).shape == (0,)
    assert vm.gate_mux(a, b, a).shape == (0,)
    assert vm.gate_not(a).shape == (0,)
    assert ctx.decrypt(sk, vm.gate_xor(a, b)).shape == (0,)


def test_uint_min_as_one_cuda_graph(ctx, key_pair):
    """The same circuit (8 x (XNOR + MUX) + MUX = 17 gates, ~60 launches / copies / fills) recorded once with
    VirtualMachine.capture and replayed as ONE graph launch: bit-identical to the eager run, and again after the
    operands have been refilled in place; replay is not slower than eager issue."""
    import time
    from nufhe_b200.operators_integer import uint_min, uintarray_to_bitarray, bitarray_to_uintarray
    sk, ck = key_pair
    vm = ctx.make_virtual_machine(ck)
    rng = numpy.random.RandomState(9)
    count = 40

    def operands():
        xs, ys = rng.randint(0, 256, count).astype(numpy.uint8), rng.randint(0, 256, count).astype(numpy.uint8)
        return xs, ys, ctx.encrypt(sk, uintarray_to_bitarray(xs)), ctx.encrypt(sk, uintarray_to_bitarray(ys))
    xs, ys, ca, cb = operands()
    answer = vm.empty_ciphertext((count, 8))
    eager = vm.empty_ciphertext((count, 8))
    uint_min(ctx.thread, ck, eager, ca, cb, perf_params=vm.perf_params)
    g = vm.capture(lambda: uint_min(ctx.thread, ck, answer, ca, cb, perf_params=vm.perf_params), reserve_batch=2 * count)
    answer.a.zero_()
    g.replay()
    assert torch.equal(answer.a, eager.a) and torch.equal(answer.b, eager.b)
    assert (bitarray_to_uintarray(ctx.decrypt(sk, answer)) == numpy.minimum(xs, ys)).all()
    xs2, ys2, ca2, cb2 = operands()                       # new data into the captured operands, replay
    for dst, src in ((ca, ca2), (cb, cb2)):
        dst.a.copy_(src.a); dst.b.copy_(src.b); dst.current_variances.copy_(src.current_variances)
    g.replay()
    assert (bitarray_to_uintarray(ctx.decrypt(sk, answer)) == numpy.minimum(xs2, ys2)).all()
    uint_min(ctx.thread, ck, eager, ca, cb, perf_params=vm.perf_params)
    assert torch.equal(answer.a, eager.a) and torch.equal(answer.b, eager.b)

    def wall(fn):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        return time.perf_counter() - t
    t_graph = min(wall(g.replay) for _ in range(3))
    t_eager = min(wall(lambda: uint_min(ctx.thread, ck, eager, ca, cb, perf_params=vm.perf_params)) for _ in range(3))
    print('uint_min x%d: graph %.2f ms, eager %.2f ms' % (count, 1e3 * t_graph, 1e3 * t_eager))
    assert t_graph < 1.05 * t_eager


def test_engine_leaves_the_callers_device_alone(nufhe):
    """Every C entry point restores the caller's current device (one process may hold one engine per GPU).  With a
    single GPU the guard is a no-op; with two, an engine on cuda:1 must not move torch's current device."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    from nufhe_b200.engine import Engine
    torch.cuda.set_device(0)
    eng1 = Engine(1)
    x = torch.zeros((4, 1024), dtype=torch.int32, device='cuda:1')
    eng1.ntt_forward_i32(x)
    assert torch.cuda.current_device() == 0
    assert torch.empty(1, device='cuda').device.index == 0


def test_multi_kernel_bootstrap_equals_fused(ctx, key_pair, nufhe):
    """`single_kernel_bootstrap=False` runs the reference's literal sequence of separate launches
    (bootstrap.py:96-229, gates.py:108-121, :629-664); the ciphertexts must equal the fused kernel's bit for bit."""
    sk, ck = key_pair
    rng = numpy.random.RandomState(5)
    bits = [rng.randint(0, 2, size=3).astype(bool) for _ in range(3)]
    cts = [ctx.encrypt(sk, b) for b in bits]
    vm1 = ctx.make_virtual_machine(ck)
    vm2 = ctx.make_virtual_machine(ck, perf_params=nufhe.PerformanceParameters(ck.params, single_kernel_bootstrap=False))
    assert vm1.perf_params.single_kernel_bootstrap and not vm2.perf_params.single_kernel_bootstrap
    for name, n_args in (('gate_nand', 2), ('gate_xor', 2), ('gate_mux', 3)):
        r1 = getattr(vm1, name)(*cts[:n_args])
        r2 = getattr(vm2, name)(*cts[:n_args])
        assert (host(r1.a) == host(r2.a)).all() and (host(r1.b) == host(r2.b)).all(), name
        assert numpy.allclose(host(r1.current_variances), host(r2.current_variances))
    assert (ctx.decrypt(sk, vm2.gate_nand(cts[0], cts[1])) == ~(bits[0] & bits[1])).all()


def test_bootstrap_seams(ctx, key_pair, okeys, nufhe):
    """The inner seams a caller of the reference can hook: bootstrap(), blind_rotate_and_extract(),
    blind_rotate(), mux_rotate() -- multi-kernel and fused give the oracle's bits."""
    from nufhe_b200.bootstrap import bootstrap, blind_rotate, mux_rotate
    from nufhe_b200.tlwe import TLweSampleArray
    from nufhe_b200.lwe import LweSampleArray
    thr = ctx.thread
    sk, ck = key_pair
    bk, ks = ck.bootstrap_key, ck.keyswitch_key
    x = ctx.encrypt(sk, numpy.array([True, False]))
    want = O.bootstrap(host(x.a), host(x.b), okeys.bk, okeys.ks)
    pp = nufhe.PerformanceParameters(ck.params, single_kernel_bootstrap=False).for_device(thr.device_params)
    for perf in (None, pp):
        res = LweSampleArray.empty(thr, ck.params.in_out_params, (2,))
        bootstrap(thr, res, bk, ks, O.MU, x, perf)
        assert (host(res.a) == want[0]).all() and (host(res.b) == want[1]).all()
    # one CMux step and a 3-row rotation against the oracle
    rng = G.rs(77)
    acc0 = G.torus32(rng, (2, 2, 1024))
    bara = G.torus32(rng, (2, 500), 0, 2048)
    accum = TLweSampleArray.empty(thr, bk.bk_params.tlwe_params, (2,))
    accum.a.coeffs.copy_(thr.to_device(acc0))
    result = TLweSampleArray.empty(thr, bk.bk_params.tlwe_params, (2,))
    mux_rotate(thr, result, accum, bk.tgsw, 4, thr.to_device(bara))
    one = O.shift_torus_polynomial(acc0, bara, 4, minus_one=True)
    one = (O.tgsw_external_mul(one, okeys.bk, 4).view(numpy.uint32) + acc0.view(numpy.uint32)).view(numpy.int32)
    assert (host(result.a.coeffs) == one).all()
    blind_rotate(thr, accum, bk, thr.to_device(bara), 3)
    assert (host(accum.a.coeffs) == O.blind_rotate(acc0, okeys.bk[:3], numpy.ascontiguousarray(bara[:, :3]))).all()


def _sha(t, unsigned=False):
    import hashlib
    return hashlib.sha256(numpy.ascontiguousarray(host(t, unsigned)).tobytes()).hexdigest()


def test_mask_size_2_against_reference_golden(nufhe, golden):
    """`NuFHEParameters(tlwe_mask_size=2)`: keys from the seed must equal the reference's (digests in
    tests/golden/k2.npz, produced by the reference's closures), and gate_nand on the golden ciphertexts must return
    the reference's bits -- 500 CMux steps of the multi-kernel path with 3 accumulator polynomials."""
    g = golden('k2')
    ctx = nufhe.Context(rng=nufhe.DeterministicRNG(int(g['seed'])))
    sk, ck = ctx.make_key_pair(tlwe_mask_size=2)
    assert ck.params == nufhe.NuFHEParameters(tlwe_mask_size=2)
    bk = ck.bootstrap_key.tgsw.samples.a.coeffs
    assert tuple(bk.shape) == (500, 3, 2, 3, 1024)
    assert _sha(sk.lwe_key.key) == str(g['lwe_key_sha'])
    assert (host(bk[0], True) == g['bk_row0']).all() and (host(bk[499], True) == g['bk_row499']).all()
    assert _sha(bk, True) == str(g['bk_sha'])
    ks = ck.keyswitch_key.lwe
    assert tuple(ks.a.shape) == (2048, 8, 4, 500)
    assert _sha(ks.a) == str(g['ks_a_sha']) and _sha(ks.b) == str(g['ks_b_sha'])
    c1, c2 = ctx.encrypt(sk, G.GATE_BITS_A[:2]), ctx.encrypt(sk, G.GATE_BITS_B[:2])
    assert (host(c1.a) == g['c1_a']).all() and (host(c1.b) == g['c1_b']).all()
    assert (host(c2.a) == g['c2_a']).all() and (host(c2.b) == g['c2_b']).all()
    vm = ctx.make_virtual_machine(ck)
    assert not vm.perf_params.single_kernel_bootstrap
    r = vm.gate_nand(c1, c2)
    assert (host(r.a) == g['nand_a']).all() and (host(r.b) == g['nand_b']).all()
    assert (ctx.decrypt(sk, r) == g['nand_bits']).all()
    with pytest.raises(ValueError):
        ctx.make_virtual_machine(ck, perf_params=nufhe.PerformanceParameters(ck.params, single_kernel_bootstrap=True))
    # the other gates decrypt to their truth tables; serialization keeps the parameters
    a = numpy.array([True, True, False, False])
    b = numpy.array([True, False, True, False])
    ca, cb = ctx.encrypt(sk, a), ctx.encrypt(sk, b)
    assert (ctx.decrypt(sk, vm.gate_xor(ca, cb)) == (a ^ b)).all()
    assert (ctx.decrypt(sk, vm.gate_mux(ca, cb, vm.gate_not(cb))) == numpy.where(a, b, ~b)).all()
    ck2 = ctx.load_cloud_key(ck.dumps())
    assert ck2 == ck and ck2.params == ck.params


def test_find_devices(nufhe):
    devs = nufhe.find_devices()
    assert len(devs) >= 1 and devs[0].api_name == 'CUDA'
    import pickle
    d = pickle.loads(pickle.dumps(devs[0]))
    c = nufhe.Context(device_id=d)
    assert c.thread.device.index == d.device_id
