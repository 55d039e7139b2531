"""Integer circuits built from gates (reference: nufhe/operators_integer.py): bit <-> uint helpers and the
`uint_min` comparator circuit (one XNOR and one MUX per bit, then a final MUX).  No kernel of its own."""
import numpy

from .api_low_level import empty_ciphertext
from .gates import gate_constant, gate_xnor, gate_mux


def uintarray_to_bitarray(xs, itemsize=None):
    """Big-endian bits of an unsigned integer array: shape xs.shape + (itemsize,) (operators_integer.py:40-45)."""
    xs = numpy.asarray(xs)
    assert numpy.issubdtype(xs.dtype, numpy.unsignedinteger)
    if itemsize is None:
        itemsize = xs.itemsize * 8
    shifts = numpy.arange(itemsize - 1, -1, -1, dtype=numpy.uint64)
    return ((xs.astype(numpy.uint64)[..., None] >> shifts) & numpy.uint64(1)).astype(bool)


def bitarray_to_uintarray(xs):
    """Inverse of uintarray_to_bitarray for item sizes 8, 16, 32, 64 (operators_integer.py:48-61)."""
    xs = numpy.asarray(xs).astype(bool)
    itemsize = xs.shape[-1]
    dtype = {8: numpy.uint8, 16: numpy.uint16, 32: numpy.uint32, 64: numpy.uint64}[itemsize]
    shifts = numpy.arange(itemsize - 1, -1, -1, dtype=numpy.uint64)
    return (xs.astype(numpy.uint64) << shifts).sum(-1, dtype=numpy.uint64).astype(dtype)


def uint_min(thread, cloud_key, answer, a, b, perf_params=None):
    """answer = elementwise min(a, b) of encrypted big-endian unsigned integers of shape (count, bits)
    (operators_integer.py:64-95).  Walks from the least significant bit keeping "is b smaller so far"."""
    params = cloud_key.params
    itemsize = answer.shape[-1]
    lead = tuple(a.shape[:-1])
    carry = empty_ciphertext(thread, params, lead + (1,))
    same = empty_ciphertext(thread, params, lead + (1,))
    gate_constant(thread, cloud_key, carry, False)
    for i in reversed(range(itemsize)):
        a_bit = a[..., i:i + 1]
        b_bit = b[..., i:i + 1]
        gate_xnor(thread, cloud_key, same, a_bit, b_bit, perf_params=perf_params)            # a_i == b_i ?
        gate_mux(thread, cloud_key, carry, same, carry, a_bit, perf_params=perf_params)      # equal: keep; else a_i
    # carry = 1 iff b < a
    gate_mux(thread, cloud_key, answer, carry, b, a, perf_params=perf_params)
