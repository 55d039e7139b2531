"""Stand-alone batched transforms behind the reference's interface (nufhe/transform/computation.py:28-99 `Transform`,
nufhe/polynomial_transform_ntt.py:120-131 `ForwardTransform` / `InverseTransform`, nufhe/polynomial_transform.py).

The reference renders a Reikna computation per (batch shape, direction, conversion); here one ahead-of-time compiled
kernel per (direction, conversion) handles any batch (`nb_ntt_forward_i32/_u64`, `nb_ntt_inverse_i32/_u64`), so
`compile(thr)` only binds the engine.  Natural order in and out, results identical to `ntt_transform_ref`."""
import numpy

N = 1024


class _NTT1024:
    """What `Transform` reads from the reference's transform module (transform/ntt.py:96-166)."""
    elem_dtype = numpy.uint64
    transform_length = N
    polynomial_length = N
    threads_per_transform = 64          # the engine's: 4 polynomials per sweep of 256 threads
    use_constant_memory = False


def ntt1024(**kwds):
    return _NTT1024()


class Transform:
    """output <- NTT (or inverse NTT) of input over the last axis; `i32_conversion` reads (forward) or writes
    (inverse) Torus32 coefficients instead of field elements.  `transforms_per_block` is accepted and ignored;
    `kernel_repetitions` repeats the launch like the reference's benchmark mode (same result)."""

    def __init__(self, transform=None, batch_shape=(), inverse=False, i32_conversion=False, transforms_per_block=4,
                 kernel_repetitions=1):
        self._transform = transform if transform is not None else _NTT1024()
        if self._transform.transform_length != N:
            raise ValueError("Only the 1024-point NTT is supported")
        self._batch_shape = tuple(batch_shape)
        self._inverse = inverse
        self._i32_conversion = i32_conversion
        self._kernel_repetitions = int(kernel_repetitions)
        self._thr = None

    def compile(self, thr):
        bound = Transform(self._transform, self._batch_shape, self._inverse, self._i32_conversion,
                          kernel_repetitions=self._kernel_repetitions)
        bound._thr = thr
        return bound

    def __call__(self, output, input_):
        thr = self._thr
        if thr is None:
            raise ValueError("Transform must be compiled for an engine first: Transform(...).compile(thr)")
        shape = self._batch_shape + (N,)
        if tuple(input_.shape) != shape or tuple(output.shape) != shape:
            raise ValueError("Transform was created for arrays of shape {s}".format(s=shape))
        for _ in range(max(1, self._kernel_repetitions)):
            if self._inverse:
                res = thr.ntt_inverse_i32(input_) if self._i32_conversion else thr.ntt_inverse_u64(input_)
            else:
                res = thr.ntt_forward_i32(input_) if self._i32_conversion else thr.ntt_forward_u64(input_)
        output.copy_(res.reshape(output.shape))
        return output


def transformed_dtype():
    return numpy.dtype('uint64')


def transformed_length(polynomial_degree):
    return polynomial_degree


def ForwardTransform(batch_shape, polynomial_degree, perf_params=None):
    assert polynomial_degree == N
    return Transform(_NTT1024(), batch_shape, i32_conversion=True)


def InverseTransform(batch_shape, polynomial_degree, perf_params=None):
    assert polynomial_degree == N
    return Transform(_NTT1024(), batch_shape, i32_conversion=True, inverse=True)


def get_transform(transform_type):
    """polynomial_transform.py:33-37; only the NTT exists here (the FFT path is out of scope)."""
    if transform_type == 'NTT':
        import sys
        return sys.modules[__name__]
    raise ValueError("transform_type " + repr(transform_type) + " is not supported by the B200 engine (NTT only)")


def transform_supported(device_params, transform_type):
    return transform_type == 'NTT'


def max_supported_transforms_per_block(device_params, transform_type):
    return 4
