"""Homomorphic gates (reference: nufhe/gates.py).

Every bootstrapped binary gate is "(0, c) + s_a * a + s_b * b, then bootstrap" (gates.py:108-121 and
its siblings); the constants below are the reference's, and the linear part is folded into the head of
the bootstrap kernel instead of being three separate launches."""
import numpy

from .numeric_functions import phase_to_t32
from .lwe import (
    LweSampleArray, lwe_negate, lwe_copy, lwe_noiseless_trivial, lwe_noiseless_trivial_constant, lwe_add_mul_to,
    lwe_add_to, lwe_sub_to, lwe_keyswitch)
from .bootstrap import bootstrap_affine, bootstrap, _single_kernel
from .tgsw import engine_format
from .performance import PerformanceParametersForDevice


def get_shape(obj):
    """Message shape of a gate argument: anything with `.shape`, or a (nested) list of bits
    (behaviour of gates.py:40-46)."""
    shape = getattr(obj, 'shape', None)
    if shape is not None:
        return tuple(shape)
    if isinstance(obj, list):
        return numpy.shape(obj)
    raise ValueError("An object of type %s is not array-like" % type(obj))


def result_shape(*shapes):
    """Shape of a gate result for arguments of the given message shapes: right-aligned, an extent of 1 (or a missing
    leading axis) stretches to the other arguments' extent, two different extents above 1 do not combine
    (what gates.py:49-68 computes pair by pair)."""
    if len(shapes) == 1:
        return tuple(shapes[0])
    rank = max(len(shape) for shape in shapes)
    extents = [1] * rank
    for shape in shapes:
        for axis, extent in enumerate(shape, rank - len(shape)):
            have = extents[axis]
            if extent > 1 and have > 1 and extent != have:
                raise ValueError("Incompatible shapes: " + ", ".join(str(tuple(s)) for s in shapes))
            if have <= 1:
                extents[axis] = extent
    return tuple(extents)


def _check_broadcast(derived, dest_shape, what):
    dest_shape = tuple(dest_shape)
    if len(derived) > len(dest_shape) or derived != dest_shape[len(dest_shape) - len(derived):]:
        raise ValueError("The shape of %s %s cannot be broadcasted to the shape of the destination %s"
                         % (what, derived, dest_shape))


def check_shape(result, *args):
    """The broadcast shape of the arguments has to be the trailing part of the destination's shape (gates.py:71-78)."""
    _check_broadcast(result_shape(*[tuple(arg.shape) for arg in args]), result.shape,
                     "the result derived from the arguments")


MU = phase_to_t32(1, 8)

# name -> (constant numerator, denominator, sign of a, sign of b); gates.py:81-597
_BINARY_GATES = {
    'nand': (1, 8, -1, -1),     # gates.py:108-115
    'or': (1, 8, 1, 1),         # :150-157
    'and': (-1, 8, 1, 1),       # :192-199
    'xor': (1, 4, 2, 2),        # :234-241
    'xnor': (-1, 4, -2, -2),    # :276-283
    'nor': (-1, 8, -1, -1),     # :415-422
    'andny': (-1, 8, -1, 1),    # :457-464
    'andyn': (-1, 8, 1, -1),    # :499-506
    'orny': (1, 8, -1, 1),      # :541-548
    'oryn': (1, 8, 1, -1),      # :583-590
}


def _binary_gate(name, thr, cloud_key, result, a, b, perf_params):
    check_shape(result, a, b)
    num, den, sa, sb = _BINARY_GATES[name]
    if not _single_kernel(perf_params, cloud_key.bootstrap_key):
        # the reference's own sequence (gates.py:108-121): trivial constant, two linear updates, bootstrap
        bk = cloud_key.bootstrap_key
        temp = LweSampleArray.empty(thr, bk.in_out_params, result.shape)
        lwe_noiseless_trivial_constant(thr, temp, phase_to_t32(num, den))
        lwe_add_mul_to(thr, temp, sa, a)
        lwe_add_mul_to(thr, temp, sb, b)
        bootstrap(thr, result, bk, cloud_key.keyswitch_key, MU, temp, perf_params)
        return
    bootstrap_affine(
        thr, result, cloud_key.bootstrap_key, cloud_key.keyswitch_key, MU,
        a, b, phase_to_t32(num, den), sa, sb)


def _make_binary(name, doc):
    def gate(thr, cloud_key, result: LweSampleArray, a: LweSampleArray, b: LweSampleArray,
             perf_params: PerformanceParametersForDevice = None):
        _binary_gate(name, thr, cloud_key, result, a, b, perf_params)
    gate.__name__ = 'gate_' + name
    gate.__doc__ = doc + """

    The shapes of ``a`` and ``b`` should be broadcastable to the shape of ``result``.

    :param thr: the engine (analogue of a ``reikna`` ``Thread``).
    :param cloud_key: the cloud key.
    :param result: an empty ciphertext where the result will be stored.
    :param a: the ciphertext with the first argument.
    :param b: the ciphertext with the second argument.
    :param perf_params: accepted for API compatibility.
    """
    return gate


gate_nand = _make_binary('nand', "Homomorphic bootstrapped NAND gate.")
gate_or = _make_binary('or', "Homomorphic bootstrapped OR gate.")
gate_and = _make_binary('and', "Homomorphic bootstrapped AND gate.")
gate_xor = _make_binary('xor', "Homomorphic bootstrapped XOR gate.")
gate_xnor = _make_binary('xnor', "Homomorphic bootstrapped XNOR gate.")
gate_nor = _make_binary('nor', "Homomorphic bootstrapped NOR gate.")
gate_andny = _make_binary('andny', "Homomorphic bootstrapped ANDNY (``(not a) and b``) gate.")
gate_andyn = _make_binary('andyn', "Homomorphic bootstrapped ANDYN (``a and (not b)``) gate.")
gate_orny = _make_binary('orny', "Homomorphic bootstrapped ORNY (``(not a) or b``) gate.")
gate_oryn = _make_binary('oryn', "Homomorphic bootstrapped ORYN (``a or (not b)``) gate.")


def gate_not(thr, cloud_key, result: LweSampleArray, a: LweSampleArray, perf_params=None):
    """Homomorphic NOT gate (does not need to be bootstrapped); gates.py:292-317."""
    check_shape(result, a)
    lwe_negate(thr, result, a)


def gate_copy(thr, cloud_key, result: LweSampleArray, a: LweSampleArray, perf_params=None):
    """Homomorphic COPY gate (does not need to be bootstrapped); gates.py:320-345."""
    check_shape(result, a)
    lwe_copy(thr, result, a)


def gate_constant(thr, cloud_key, result: LweSampleArray, vals, perf_params=None):
    """Homomorphic CONSTANT gate: trivial encryptions of the given bits; gates.py:348-387."""
    vals = numpy.asarray(vals)
    # the same rule as for ciphertext arguments (gates.py:371: check_shape(result, vals)): size-1 axes broadcast
    _check_broadcast(result_shape(tuple(vals.shape)), result.shape, "the values")
    mus = numpy.where(vals.astype(bool), MU, -MU).astype(numpy.int32)
    if mus.ndim == 0:
        # one bit for the whole array (gates.py:384-385): a fill, no host -> device copy (safe under graph capture)
        lwe_noiseless_trivial_constant(thr, result, int(mus))
    else:
        lwe_noiseless_trivial(thr, result, thr.to_device(numpy.ascontiguousarray(mus)))


def gate_mux(thr, cloud_key, result: LweSampleArray, a: LweSampleArray, b: LweSampleArray,
             c: LweSampleArray, perf_params=None):
    """Homomorphic bootstrapped MUX (``b if a else c``) gate; gates.py:600-664.
    Two blind rotations without key switch, then ONE key switch of ``(0,1/8) + u1 + u2`` with the
    sum folded into the key-switch kernel's load."""
    check_shape(result, a, b, c)
    bk, ks = cloud_key.bootstrap_key, cloud_key.keyswitch_key
    and_const = phase_to_t32(-1, 8)
    shape = tuple(result.shape)
    if not _single_kernel(perf_params, bk):
        # the reference's own sequence (gates.py:629-664)
        in_out, extracted = bk.in_out_params, bk.extract_params
        temp = LweSampleArray.empty(thr, in_out, shape)
        temp1 = LweSampleArray.empty(thr, extracted, shape)
        u1 = LweSampleArray.empty(thr, extracted, shape)
        u2 = LweSampleArray.empty(thr, extracted, shape)
        lwe_noiseless_trivial_constant(thr, temp, and_const)       # AND(a, b), no key switch
        lwe_add_to(thr, temp, a)
        lwe_add_to(thr, temp, b)
        bootstrap(thr, u1, bk, ks, MU, temp, perf_params, no_keyswitch=True)
        lwe_noiseless_trivial_constant(thr, temp, and_const)       # AND(not a, c), no key switch
        lwe_sub_to(thr, temp, a)
        lwe_add_to(thr, temp, c)
        bootstrap(thr, u2, bk, ks, MU, temp, perf_params, no_keyswitch=True)
        lwe_noiseless_trivial_constant(thr, temp1, phase_to_t32(1, 8))
        lwe_add_to(thr, temp1, u1)
        lwe_add_to(thr, temp1, u2)
        lwe_keyswitch(thr, result, ks, temp1)
        return

    def expand(x):
        a, b = x.a, x.b
        if tuple(b.shape) != shape:
            while b.dim() < len(shape):
                a, b = a.unsqueeze(0), b.unsqueeze(0)
            a, b = a.expand(shape + (a.shape[-1],)), b.expand(shape)
        return (a.contiguous(), b.contiguous())

    pa, pb, pc = expand(a), expand(b), expand(c)
    # AND(a, b) and AND(not a, c) as two jobs of ONE launch, then one key switch of (0,1/8) + u1 + u2
    u1, u2 = thr.bootstrap_extract2((pa, pb, and_const, 1, 1), (pa, pc, and_const, -1, 1), MU,
                                    engine_format(thr, bk.tgsw))
    dense = result.a.is_contiguous() and result.b.is_contiguous()
    res_a, res_b, res_cv = thr.keyswitch(ks.device_arrays(), u1, u2, c=phase_to_t32(1, 8),
                                         out=(result.a, result.b) if dense else None, want_cv=True)
    if not dense:
        result.a.copy_(res_a.reshape(result.a.shape))
        result.b.copy_(res_b.reshape(result.b.shape))
    result.current_variances.copy_(res_cv.reshape(result.current_variances.shape))
