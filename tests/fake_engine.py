"""A CPU test double for nufhe_b200.engine.Engine, backed by the oracle.  TEST INFRASTRUCTURE ONLY.

The product has no CPU path: `Engine()` raises without CUDA and nothing under nufhe_b200/ imports this module (or the
oracle).  The double lets the CPU-only suite drive the real Python host layer -- key generation in the reference's RNG
order, gates and their broadcasting, the multi-kernel bootstrap loop, the k = 2 flow, serialization -- and compare it
with the reference's golden vectors, so that a mistake in that layer is caught without a GPU.  Every method has the
signature and the semantics of the Engine method of the same name (nufhe_b200/engine.py) on CPU torch tensors."""
import numpy
import torch

from oracle import oracle as O

N = 1024


def _np(t, unsigned=False):
    a = t.detach().cpu().contiguous().numpy()
    return a.view(numpy.uint64) if unsigned else a


def _t(arr):
    arr = numpy.ascontiguousarray(arr)
    if arr.dtype == numpy.uint64:
        arr = arr.view(numpy.int64)
    return torch.from_numpy(arr)


def _wrap32(x):
    return (x.astype(numpy.int64) & 0xffffffff).astype(numpy.uint32).view(numpy.int32)


class _DeviceParams:
    compute_units = 1
    max_work_group_size = 1024
    local_mem_size = 227 * 1024
    name = 'oracle-backed test double'


class FakeEngine:
    SHIFT_INVERT, SHIFT_MINUS_ONE, SHIFT_PLAIN = 0, 1, 2

    def __init__(self):
        self.device = torch.device('cpu')
        self.device_params = _DeviceParams()
        self.calls = {}

    def _count(self, name):
        self.calls[name] = self.calls.get(name, 0) + 1

    def synchronize(self):
        pass

    def build_info(self):
        return 'fake engine (oracle)'

    def empty(self, shape, dtype):
        return torch.empty(tuple(shape), dtype=dtype)

    def to_device(self, arr):
        return _t(arr)

    @staticmethod
    def to_host(t, unsigned=False):
        return _np(t, unsigned)

    # --- transforms and field ops
    def ntt_forward_i32(self, x):
        return _t(O.ntt_forward_i32(_np(x).reshape(-1, N))).reshape(x.shape)

    def ntt_forward_u64(self, x):
        return _t(O.ntt_forward_u64(_np(x, True).reshape(-1, N))).reshape(x.shape)

    def ntt_inverse_i32(self, x):
        return _t(O.ntt_inverse_i32(_np(x, True).reshape(-1, N))).reshape(x.shape)

    def ntt_inverse_u64(self, x):
        return _t(O.ntt_inverse_u64(_np(x, True).reshape(-1, N))).reshape(x.shape)

    def ff_op(self, op, a, b=None):
        from nufhe_b200 import _native as nv
        an = _np(a, True).ravel() % numpy.uint64(O.P)
        if b is not None:
            bn = _np(b, True).ravel()
            if bn.size != an.size:
                bn = numpy.tile(bn, an.size // bn.size)
        if op == nv.FF_PREPARE:
            res = O.ff_prepare_for_mul(an)
        elif op == nv.FF_LSH or op == nv.FF_LSH_CONST:
            res = O.ff_lsh(an, (bn % 192).astype(numpy.uint32))
        else:
            fn = {nv.FF_ADD: O.ff_add, nv.FF_SUB: O.ff_sub, nv.FF_MUL: O.ff_mul, nv.FF_MUL_PREPARED: O.ff_mul_prepared}[op]
            res = fn(an, bn % numpy.uint64(O.P))
        return _t(res).reshape(a.shape)

    # --- the fused path (the double keeps the key in the reference's layout)
    def bk_prepare(self, bk_ref):
        return bk_ref

    def external_product(self, accum, bk_int, row):
        self._count('external_product')
        res = O.tgsw_external_mul(_np(accum).reshape(-1, 2, N), _np(bk_int, True), row)
        accum.copy_(_t(res).reshape(accum.shape))
        return accum

    def blind_rotate(self, accum, bara, bk_int, extract=True, return_accum=False):
        acc = _np(accum).reshape(-1, 2, N)
        B = acc.shape[0]
        rot = O.blind_rotate(acc, _np(bk_int, True), _np(bara).reshape(B, -1))
        out_a = out_b = None
        if extract:
            ea, eb = O.tlwe_extract_lwe_samples(rot)
            out_a, out_b = _t(ea), _t(eb)
        return out_a, out_b, (_t(rot).reshape(accum.shape) if return_accum else None)

    def _affine(self, x1, x2, c, s1, s2):
        a = _np(x1[0]).astype(numpy.int64) * s1
        b = _np(x1[1]).astype(numpy.int64) * s1 + int(c)
        if x2 is not None:
            a = a + _np(x2[0]).astype(numpy.int64) * s2
            b = b + _np(x2[1]).astype(numpy.int64) * s2
        return _wrap32(a), _wrap32(b)

    def bootstrap_extract(self, x1, x2, c, s1, s2, mu, bk_int, out=None):
        self._count('bootstrap_extract')
        a, b = self._affine(x1, x2, c, s1, s2)
        B = b.size
        ea, eb = O.bootstrap(a.reshape(B, -1), b.reshape(B), _np(bk_int, True), None, mu)
        ea, eb = _t(ea), _t(eb)
        if out is not None:
            out[0].copy_(ea.reshape(out[0].shape))
            out[1].copy_(eb.reshape(out[1].shape))
            return out
        return ea, eb

    def bootstrap_extract2(self, job_a, job_b, mu, bk_int):
        return (self.bootstrap_extract(job_a[0], job_a[1], job_a[2], job_a[3], job_a[4], mu, bk_int),
                self.bootstrap_extract(job_b[0], job_b[1], job_b[2], job_b[3], job_b[4], mu, bk_int))

    def keyswitch(self, ks, src1, src2=None, c=0, out=None, want_cv=False):
        self._count('keyswitch')
        a, b = self._affine(src1, src2, c, 1, 1)
        shape = tuple(src1[1].shape)
        ra, rb, rcv = O.lwe_keyswitch(_np(ks[0]), _np(ks[1]), _np(ks[2]), a.reshape(b.size, -1), b.reshape(-1))
        ra, rb, rcv = _t(ra).reshape(shape + (ra.shape[-1],)), _t(rb).reshape(shape), _t(rcv).reshape(shape)
        if out is not None:
            out[0].copy_(ra.reshape(out[0].shape))
            out[1].copy_(rb.reshape(out[1].shape))
            ra, rb = out
        return ra, rb, (rcv if want_cv else None)

    def lwe_dot(self, a, key, add1=None, add2=None, sign=1, out=None):
        """nb_lwe_dot restated with NumPy (vec_mul_mat of nufhe/lwe_cpu.py:22-23 + the addends)."""
        an, kn = _np(a).astype(numpy.int64), _np(key).astype(numpy.int64)
        v = int(sign) * (an * kn).sum(-1)
        for add in (add1, add2):
            if add is not None:
                v = v + _np(add).astype(numpy.int64).reshape(v.shape)
        res = _t(_wrap32(v)).reshape(tuple(a.shape[:-1]))
        if out is not None:
            out.copy_(res)
            return out
        return res

    def make_keyswitch_key(self, ks_a, ks_b, ks_cv, in_key, out_key, noises_a, noises_b, log2_base, noise_variance):
        """nb_make_keyswitch_key restated with NumPy (MakeLweKeyswitchKeyReference, nufhe/lwe_cpu.py:26-59)."""
        in_size, t, base, n = ks_a.shape
        na, nb_ = _np(noises_a).astype(numpy.int64), _np(noises_b).astype(numpy.int64)
        hs = numpy.arange(1, base, dtype=numpy.int64)[None, None, :]
        js = numpy.arange(t, dtype=numpy.int64)[None, :, None]
        messages = _np(in_key).astype(numpy.int64)[:, None, None] * hs * (2 ** (32 - (js + 1) * log2_base))
        b = messages + nb_ + (na * _np(out_key).astype(numpy.int64)).sum(-1)
        ks_a[:, :, 0, :] = 0
        ks_a[:, :, 1:, :] = noises_a
        ks_b[:, :, 0] = 0
        ks_b[:, :, 1:] = _t(_wrap32(b))
        ks_cv[:, :, 0] = 0
        ks_cv[:, :, 1:] = float(noise_variance)

    def lwe_affine(self, res, x1, x2, c, s1, s2):
        zero = (torch.zeros_like(res[0]), torch.zeros_like(res[1]))
        a, b = self._affine(x1 if x1 is not None else zero, x2, c, s1 if x1 is not None else 0, s2)
        res[0].copy_(_t(a).reshape(res[0].shape))
        res[1].copy_(_t(b).reshape(res[1].shape))
        return res

    # --- the separate steps of the multi-kernel path
    def shift_torus_polynomial(self, result, source, powers, power_idx=0, polys_per_power=1, mode=0):
        self._count('shift_torus_polynomial')
        n = result.shape[-1]
        assert n == N
        src = _np(source).reshape(-1, polys_per_power, n)
        pw = _np(powers).reshape(src.shape[0], -1)
        res = O.shift_torus_polynomial(src, pw if pw.shape[1] > 1 else pw[:, 0],
                                       power_idx if pw.shape[1] > 1 else None,
                                       minus_one=(mode == self.SHIFT_MINUS_ONE), invert_powers=(mode == self.SHIFT_INVERT))
        result.copy_(_t(res).reshape(result.shape))
        return result

    def tlwe_noiseless_trivial(self, acc, cv, mu):
        acc.zero_()
        acc[..., acc.shape[-2] - 1, :] = mu
        if cv is not None:
            cv.zero_()
        return acc

    def tlwe_extract_lwe_samples(self, out_a, out_b, acc):
        a = _np(acc)
        k = a.shape[-2] - 1
        a = a.reshape(-1, k + 1, N)
        res = numpy.empty((a.shape[0], k, N), numpy.int32)
        res[:, :, 0] = a[:, :k, 0]
        res[:, :, 1:] = _wrap32(-a[:, :k, :0:-1].astype(numpy.int64))
        out_a.copy_(_t(res).reshape(out_a.shape))
        out_b.copy_(_t(numpy.ascontiguousarray(a[:, k, 0])).reshape(out_b.shape))

    def tlwe_add_to(self, res, src, res_cv=None, src_cv=None):
        self._count('tlwe_add_to')
        res.copy_(_t(_wrap32(_np(res).astype(numpy.int64) + _np(src).astype(numpy.int64))).reshape(res.shape))
        if res_cv is not None:
            res_cv.add_(src_cv)

    def tgsw_decompose(self, acc, decomp_length, bs_log2_base, offset):
        assert decomp_length == 2 and bs_log2_base == 10 and int(offset) == -2145386496
        return _t(O.tgsw_decompose_k(_np(acc)))

    def tgsw_mac(self, tr, bk_row, mask_size, decomp_length):
        self._count('tgsw_mac')
        return _t(O.tgsw_mac_k(_np(tr, True).reshape(-1, mask_size + 1, decomp_length, N), _np(bk_row, True)))

    def t32_to_phase(self, out, messages, mspace_size):
        out.copy_(_t(O.t32_to_phase(_np(messages), mspace_size)).reshape(out.shape))
        return out
