"""Single-kernel blind rotation entry point (reference: nufhe/blind_rotate.py:262-281)."""
import torch

from .lwe import LweSampleArray, _keyswitch_into
from .tgsw import engine_format


def BlindRotate_gpu(lwe_out: LweSampleArray, accum, bk, ks, bara, perf_params=None, no_keyswitch=False, thr=None):
    """lwe_out <- keyswitch(extract(blind_rotate(accum, bara)))  (keyswitch skipped if no_keyswitch).
    `accum` is a TLweSampleArray with coefficients of shape (..., 2, 1024); `bara` an int32 tensor
    (..., n) of rotation amounts in [0, 2N).  `thr` is the Engine (the reference finds the Thread
    through the arrays; torch tensors do not carry one)."""
    if thr is None:
        raise ValueError("BlindRotate_gpu needs the engine: pass thr=")
    bk_int = engine_format(thr, bk.tgsw)
    coeffs = accum.a.coeffs
    out_a, out_b, _ = thr.blind_rotate(coeffs, bara, bk_int)
    shape = tuple(lwe_out.shape)
    if no_keyswitch:
        lwe_out.a.copy_(out_a.reshape(lwe_out.a.shape))
        lwe_out.b.copy_(out_b.reshape(lwe_out.b.shape))
    else:
        ext = LweSampleArray(
            bk.extract_params, out_a.reshape(shape + (out_a.shape[-1],)), out_b.reshape(shape),
            torch.zeros(shape, dtype=torch.float32, device=out_b.device))
        _keyswitch_into(thr, lwe_out, ks, ext, None, 0)
