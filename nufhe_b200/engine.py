"""Engine: one (CUDA device, stream) binding of the native library -- what a Reikna `Thread` is to the
reference (nufhe/api_high_level.py:153-181).  All methods take torch CUDA tensors (int32 for Torus32
data, int64 holding the uint64 bit patterns of field elements) and enqueue work on the engine's stream.
PyTorch is used for device memory and streams only.
"""
import ctypes

import numpy
import torch

from . import _native

N = 1024


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class Engine:

    def __init__(self, device=None, stream=None):
        if not torch.cuda.is_available():
            raise RuntimeError('nufhe_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
        self.lib = _native.load()
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device('cuda', device if isinstance(device, int) else torch.device(device).index or 0)
        self.torch_stream = stream
        handle = ctypes.c_void_p()
        stream_ptr = ctypes.c_void_p(stream.cuda_stream if stream is not None else 0)
        rc = self.lib.nb_ctx_create(self.device.index, stream_ptr, ctypes.byref(handle))
        self.handle = handle
        self._stream_ptr = stream_ptr.value or 0
        if rc != _native.NB_OK:
            try:
                _native.check(handle, rc, 'nb_ctx_create')
            finally:
                if handle:
                    self.lib.nb_ctx_destroy(handle)
                self.handle = None
        self.device_params = DeviceParams(self.device)

    def __del__(self):
        h = getattr(self, 'handle', None)
        if h:
            self.lib.nb_ctx_destroy(h)
            self.handle = None

    # --- plumbing -------------------------------------------------------------------------
    def _call(self, name, *args):
        # work goes to torch's current stream on the engine's device unless the engine was given its own stream: that
        # is what keeps the native launches ordered with torch's copies / fills and lets torch.cuda.graph capture them
        if self.torch_stream is None:
            cur = torch.cuda.current_stream(self.device).cuda_stream
            if cur != self._stream_ptr:
                _native.check(self.handle, self.lib.nb_ctx_set_stream(self.handle, ctypes.c_void_p(cur)), 'nb_ctx_set_stream')
                self._stream_ptr = cur
        _native.check(self.handle, getattr(self.lib, name)(self.handle, *args), name)

    def reserve(self, batch):
        """Pre-allocate the native scratch for launches of up to `batch` ciphertexts (needed before graph capture)."""
        self._call('nb_ctx_reserve', int(batch))

    def synchronize(self):
        self._call('nb_ctx_synchronize')

    def build_info(self):
        return self.lib.nb_build_info().decode()

    def empty(self, shape, dtype):
        return torch.empty(tuple(shape), dtype=dtype, device=self.device)

    def to_device(self, arr):
        arr = numpy.ascontiguousarray(arr)
        if arr.dtype == numpy.uint64:
            arr = arr.view(numpy.int64)
        return torch.from_numpy(arr).to(self.device)

    @staticmethod
    def to_host(t, unsigned=False):
        arr = t.detach().cpu().numpy()
        return arr.view(numpy.uint64) if unsigned else arr

    def _dense(self, t, dtype):
        assert t.is_cuda and t.dtype == dtype, (t.device, t.dtype, dtype)
        return t if t.is_contiguous() else t.contiguous()

    # --- transforms -----------------------------------------------------------------------
    def ntt_forward_i32(self, x):
        x = self._dense(x, torch.int32)
        out = self.empty(x.shape, torch.int64)
        self._call('nb_ntt_forward_i32', _ptr(x), _ptr(out), x.numel() // N)
        return out

    def ntt_forward_u64(self, x):
        x = self._dense(x, torch.int64)
        out = self.empty(x.shape, torch.int64)
        self._call('nb_ntt_forward_u64', _ptr(x), _ptr(out), x.numel() // N)
        return out

    def ntt_inverse_i32(self, x):
        x = self._dense(x, torch.int64)
        out = self.empty(x.shape, torch.int32)
        self._call('nb_ntt_inverse_i32', _ptr(x), _ptr(out), x.numel() // N)
        return out

    def ntt_inverse_u64(self, x):
        x = self._dense(x, torch.int64)
        out = self.empty(x.shape, torch.int64)
        self._call('nb_ntt_inverse_u64', _ptr(x), _ptr(out), x.numel() // N)
        return out

    def ff_op(self, op, a, b=None):
        a = self._dense(a, torch.int64)
        out = torch.empty_like(a)
        period = 0
        if b is not None:
            b = self._dense(b, torch.int64)
            if b.numel() != a.numel():
                assert a.numel() % b.numel() == 0
                period = b.numel()
        self._call('nb_ff_elementwise', op, _ptr(a), _ptr(b), _ptr(out), a.numel(), period)
        return out

    # --- bootstrap path -------------------------------------------------------------------
    def bk_prepare(self, bk_ref):
        bk_ref = self._dense(bk_ref, torch.int64)
        rows = bk_ref.numel() // (8 * N)
        out = self.empty((rows, self.lib.nb_bk_row_u64()), torch.int64)
        self._call('nb_bk_prepare', _ptr(bk_ref), _ptr(out), rows)
        return out

    def external_product(self, accum, bk_int, row):
        """accum (B,2,1024) int32 is updated in place: accum <- bk[row] (x) accum"""
        assert accum.is_contiguous() and accum.dtype == torch.int32
        self._call('nb_external_product', _ptr(accum), _ptr(bk_int), row, accum.numel() // (2 * N))
        return accum

    def blind_rotate(self, accum, bara, bk_int, extract=True, return_accum=False):
        accum = self._dense(accum, torch.int32)
        bara = self._dense(bara, torch.int32)
        B = accum.numel() // (2 * N)
        n = bara.shape[-1]                    # (an empty batch is a no-op in the library)
        out_a = self.empty((B, N), torch.int32) if extract else None
        out_b = self.empty((B,), torch.int32) if extract else None
        acc_out = torch.empty_like(accum) if return_accum else None
        self._call('nb_blind_rotate', _ptr(accum), _ptr(bara), _ptr(bk_int), n, _ptr(out_a), _ptr(out_b),
                   _ptr(acc_out), B)
        return out_a, out_b, acc_out

    def bootstrap_extract(self, x1, x2, c, s1, s2, mu, bk_int, out=None):
        """x = (0,c) + s1*x1 + s2*x2 (x2 may be None); returns the extracted sample (a (B,1024), b (B,))."""
        a1, b1 = x1
        a1 = self._dense(a1, torch.int32)
        b1 = self._dense(b1, torch.int32)
        a2 = b2 = None
        if x2 is not None:
            a2 = self._dense(x2[0], torch.int32)
            b2 = self._dense(x2[1], torch.int32)
        B = b1.numel()
        n = a1.shape[-1]
        out_a, out_b = out if out is not None else (self.empty((B, N), torch.int32), self.empty((B,), torch.int32))
        self._call('nb_bootstrap_extract', _ptr(a1), _ptr(b1), _ptr(a2), _ptr(b2), int(c), int(s1), int(s2),
                   int(mu), _ptr(bk_int), n, _ptr(out_a), _ptr(out_b), B)
        return out_a, out_b

    def bootstrap_extract2(self, job_a, job_b, mu, bk_int):
        """Two bootstraps in one launch.  job = (x1, x2, c, s1, s2) as for bootstrap_extract.
        Returns two extracted samples (views of one (2B, 1024) / (2B,) allocation)."""
        def parts(job):
            x1, x2, c, s1, s2 = job
            a1, b1 = self._dense(x1[0], torch.int32), self._dense(x1[1], torch.int32)
            a2 = b2 = None
            if x2 is not None:
                a2, b2 = self._dense(x2[0], torch.int32), self._dense(x2[1], torch.int32)
            return a1, b1, a2, b2, int(c), int(s1), int(s2)
        pa, pb = parts(job_a), parts(job_b)
        B = pa[1].numel()
        n = pa[0].shape[-1]
        out_a, out_b = self.empty((2 * B, N), torch.int32), self.empty((2 * B,), torch.int32)
        self._call('nb_bootstrap_extract2', _ptr(pa[0]), _ptr(pa[1]), _ptr(pa[2]), _ptr(pa[3]), pa[4], pa[5], pa[6],
                   _ptr(pb[0]), _ptr(pb[1]), _ptr(pb[2]), _ptr(pb[3]), pb[4], pb[5], pb[6], int(mu), _ptr(bk_int), n,
                   _ptr(out_a), _ptr(out_b), B)
        return (out_a[:B], out_b[:B]), (out_a[B:], out_b[B:])

    def keyswitch(self, ks, src1, src2=None, c=0, out=None, want_cv=False):
        ks_a, ks_b, ks_cv = ks
        in_size, t, base, n = ks_a.shape
        a1 = self._dense(src1[0], torch.int32)
        b1 = self._dense(src1[1], torch.int32)
        a2 = b2 = None
        if src2 is not None:
            a2 = self._dense(src2[0], torch.int32)
            b2 = self._dense(src2[1], torch.int32)
        B = b1.numel()
        if out is None:
            res_a = self.empty(tuple(b1.shape) + (n,), torch.int32)
            res_b = self.empty(tuple(b1.shape), torch.int32)
        else:
            res_a, res_b = out
        res_cv = self.empty(tuple(b1.shape), torch.float32) if want_cv else None
        self._call('nb_keyswitch', _ptr(a1), _ptr(b1), _ptr(a2), _ptr(b2), int(c), _ptr(ks_a), _ptr(ks_b),
                   _ptr(ks_cv), in_size, n, t, int(base).bit_length() - 1, _ptr(res_a), _ptr(res_b),
                   _ptr(res_cv), B)
        return res_a, res_b, res_cv

    # --- the separate steps of the multi-kernel bootstrap (bootstrap.py:96-196) -------------------
    SHIFT_INVERT, SHIFT_MINUS_ONE, SHIFT_PLAIN = 0, 1, 2

    def shift_torus_polynomial(self, result, source, powers, power_idx=0, polys_per_power=1, mode=0):
        """result (polys, N) <- X^e * source; one power per `polys_per_power` polynomials, column `power_idx` of
        `powers` (rows of `powers.shape[-1]` entries when it is 2-D per group, else one entry per group)."""
        assert result.is_contiguous() and result.dtype == torch.int32
        source = self._dense(source, torch.int32)
        powers = self._dense(powers, torch.int32)
        n_poly = result.shape[-1]
        polys = result.numel() // n_poly
        groups = polys // polys_per_power
        stride = powers.numel() // groups
        self._call('nb_shift_torus_polynomial', _ptr(result), _ptr(source), _ptr(powers), stride, power_idx,
                   polys_per_power, mode, n_poly.bit_length() - 1, polys)
        return result

    def tlwe_noiseless_trivial(self, acc, cv, mu):
        assert acc.is_contiguous() and acc.dtype == torch.int32
        mu = self._dense(mu, torch.int32)
        n_poly, k1 = acc.shape[-1], acc.shape[-2]
        self._call('nb_tlwe_noiseless_trivial', _ptr(acc), _ptr(cv), _ptr(mu), k1 - 1, n_poly.bit_length() - 1,
                   acc.numel() // (k1 * n_poly))
        return acc

    def tlwe_extract_lwe_samples(self, out_a, out_b, acc):
        assert out_a.is_contiguous() and out_b.is_contiguous()
        acc = self._dense(acc, torch.int32)
        n_poly, k1 = acc.shape[-1], acc.shape[-2]
        self._call('nb_tlwe_extract_lwe_samples', _ptr(out_a), _ptr(out_b), _ptr(acc), k1 - 1,
                   n_poly.bit_length() - 1, acc.numel() // (k1 * n_poly))

    def tlwe_add_to(self, res, src, res_cv=None, src_cv=None):
        assert res.is_contiguous() and res.dtype == torch.int32
        src = self._dense(src, torch.int32)
        if res_cv is not None:
            assert res_cv.is_contiguous()
            src_cv = self._dense(src_cv, torch.float32)
        self._call('nb_tlwe_add_to', _ptr(res), _ptr(src), res.numel(), _ptr(res_cv), _ptr(src_cv),
                   0 if res_cv is None else res_cv.numel())

    def tgsw_decompose(self, acc, decomp_length, bs_log2_base, offset):
        """acc (..., N) int32 -> (..., l, N) digits (nb_tgsw_decompose)."""
        acc = self._dense(acc, torch.int32)
        n_poly = acc.shape[-1]
        out = self.empty(tuple(acc.shape[:-1]) + (decomp_length, n_poly), torch.int32)
        self._call('nb_tgsw_decompose', _ptr(out), _ptr(acc), acc.numel() // n_poly, decomp_length, bs_log2_base,
                   int(offset), n_poly.bit_length() - 1)
        return out

    def tgsw_mac(self, tr, bk_row, mask_size, decomp_length):
        """tr (B, k+1, l, 1024) u64, bk_row (k+1, l, k+1, 1024) u64 (reference layout) -> (B, k+1, 1024) u64."""
        tr = self._dense(tr, torch.int64)
        bk_row = self._dense(bk_row, torch.int64)
        batch = tr.numel() // ((mask_size + 1) * decomp_length * N)
        out = self.empty((batch, mask_size + 1, N), torch.int64)
        self._call('nb_tgsw_mac', _ptr(out), _ptr(tr), _ptr(bk_row), batch, mask_size, decomp_length)
        return out

    def t32_to_phase(self, out, messages, mspace_size):
        assert out.is_contiguous() and out.dtype == torch.int32
        messages = self._dense(messages, torch.int32)
        self._call('nb_t32_to_phase', _ptr(out), _ptr(messages), out.numel(), int(mspace_size))
        return out

    def lwe_dot(self, a, key, add1=None, add2=None, sign=1, out=None):
        """out[i] = add1[i] + add2[i] + sign * <a[i], key> in Torus32 (nb_lwe_dot); a (..., n), key (n,)."""
        a = self._dense(a, torch.int32)
        key = self._dense(key, torch.int32)
        n = key.numel()
        shape = tuple(a.shape[:-1])
        if out is None:
            out = self.empty(shape, torch.int32)
        assert out.is_contiguous() and out.dtype == torch.int32
        add1 = self._dense(add1, torch.int32) if add1 is not None else None
        add2 = self._dense(add2, torch.int32) if add2 is not None else None
        self._call('nb_lwe_dot', _ptr(out), _ptr(a), _ptr(key), _ptr(add1), _ptr(add2), int(sign), a.numel() // n, n)
        return out

    def make_keyswitch_key(self, ks_a, ks_b, ks_cv, in_key, out_key, noises_a, noises_b, log2_base, noise_variance):
        """Fill a key-switch key (nb_make_keyswitch_key); ks_a (in, t, base, n) etc. dense, updated in place."""
        assert ks_a.is_contiguous() and ks_b.is_contiguous() and ks_cv.is_contiguous()
        in_size, t, base, n = ks_a.shape
        self._call('nb_make_keyswitch_key', _ptr(ks_a), _ptr(ks_b), _ptr(ks_cv), _ptr(self._dense(in_key, torch.int32)),
                   _ptr(self._dense(out_key, torch.int32)), _ptr(self._dense(noises_a, torch.int32)),
                   _ptr(self._dense(noises_b, torch.int32)), in_size, n, t, int(log2_base),
                   ctypes.c_float(float(noise_variance)))

    def lwe_affine(self, res, x1, x2, c, s1, s2):
        res_a, res_b = res
        B = res_b.numel()
        n = res_a.shape[-1]
        a1 = b1 = a2 = b2 = None
        if x1 is not None:
            a1, b1 = self._dense(x1[0], torch.int32), self._dense(x1[1], torch.int32)
        if x2 is not None:
            a2, b2 = self._dense(x2[0], torch.int32), self._dense(x2[1], torch.int32)
        self._call('nb_lwe_affine', _ptr(res_a), _ptr(res_b), _ptr(a1), _ptr(b1), _ptr(a2), _ptr(b2),
                   int(c), int(s1), int(s2), B, n)
        return res


class DeviceParams:
    """The few fields of Reikna's device_params that PerformanceParameters.for_device() looks at
    (nufhe/performance.py:137-236)."""

    def __init__(self, device):
        props = torch.cuda.get_device_properties(device)
        self.compute_units = props.multi_processor_count
        self.max_work_group_size = 1024
        self.local_mem_size = 227 * 1024
        self.name = props.name
