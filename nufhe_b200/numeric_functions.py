"""Scalar types and encodings (reference: nufhe/numeric_functions.py, numeric_functions_gpu.py:30-36)."""
import numpy
import torch

Torus32 = numpy.int32        # an element of R/Z stored as a wrapping int32 (value / 2^32)
Int32 = numpy.int32
ErrorFloat = numpy.float32

TORCH_DTYPES = {numpy.dtype('int32'): torch.int32, numpy.dtype('float32'): torch.float32,
                numpy.dtype('int64'): torch.int64}


def phase_to_t32(phase: int, mspace_size: int):
    """numeric_functions.py:30-31, with the int32 wrap made explicit (the reference relied on
    NumPy-1 silent overflow for e.g. phase_to_t32(-1, 8))."""
    v = (phase % mspace_size) * (2**32 // mspace_size)
    return Torus32(v - 2**32 if v >= 2**31 else v)


def double_to_t32(d):
    """numeric_functions.py:39-40"""
    return ((d - numpy.trunc(d)) * 2**32).astype(Torus32)


def t32_to_phase(thr, result, messages, mspace_size: int):
    """Mod-switch (numeric_functions.py:34-36; kernel numeric_functions_gpu.py:39-77).
    On the single-kernel gate path this is fused into the bootstrap kernel; this is the separate step of the
    multi-kernel path (nb_t32_to_phase)."""
    if result.is_contiguous():
        thr.t32_to_phase(result, messages, mspace_size)
    else:
        tmp = torch.empty(tuple(result.shape), dtype=torch.int32, device=result.device)
        thr.t32_to_phase(tmp, messages, mspace_size)
        result.copy_(tmp)
