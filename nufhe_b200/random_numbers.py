"""Host-side random number generators with the reference's draw semantics
(nufhe/random_numbers.py:46-151).  All randomness is generated on the host and uploaded, exactly as
in the reference, so a seed reproduces nufhe's keys and ciphertexts bit for bit."""
from os import urandom

import numpy

from .numeric_functions import double_to_t32, Torus32, Int32


class DeterministicRNG:
    """A fast, seedable, not cryptographically secure RNG (random_numbers.py:46-62)."""

    def __init__(self, seed=None):
        self.rng = numpy.random.RandomState(seed)

    def uniform_bool(self, shape):
        return self.rng.randint(0, 2, size=shape, dtype=Int32)

    def uniform_torus32(self, shape):
        return self.rng.randint(-2**31, 2**31, size=shape, dtype=Torus32)

    def gauss(self, shape, std_dev):
        return self.rng.normal(size=shape, scale=std_dev)


class SecureRNG:
    """Randomness for keys and encryption noise taken from the operating system's CSPRNG (`os.urandom`); the
    counterpart of the reference's `SecureRNG` (random_numbers.py:65-130), same three methods.  Nothing here has to
    reproduce the reference's byte consumption (there is no seed), so the sampling is our own: bits are unpacked from
    whole bytes, Torus32 values are four bytes each, and Gaussians come from Marsaglia's polar method on 53-bit
    uniforms, drawn in vectorised rounds until enough pairs have been accepted."""

    _MANTISSA_BITS = 53

    @staticmethod
    def _count(shape):
        return int(numpy.prod(shape, dtype=numpy.int64)) if numpy.ndim(shape) else int(shape)

    @staticmethod
    def _shape(shape):
        return tuple(shape) if numpy.ndim(shape) else (int(shape),)

    def uniform_bool(self, shape):
        count = self._count(shape)
        packed = numpy.frombuffer(urandom((count + 7) // 8), numpy.uint8)
        return numpy.unpackbits(packed, count=count).astype(Int32).reshape(self._shape(shape))

    def uniform_torus32(self, shape):
        count = self._count(shape)
        return numpy.frombuffer(urandom(4 * count), Torus32).reshape(self._shape(shape)).copy()

    def _open_unit_interval(self, count):
        """`count` doubles uniform on the open interval (0, 1): the centres of 2^53 equal cells, so 0 and 1 never occur."""
        raw = numpy.frombuffer(urandom(8 * count), numpy.uint64) >> numpy.uint64(64 - self._MANTISSA_BITS)
        return (raw.astype(numpy.float64) + 0.5) * 2.0**-self._MANTISSA_BITS

    def gauss(self, shape, std_dev):
        count = self._count(shape)
        out = numpy.empty(count + (count & 1), numpy.float64)
        filled = 0
        while filled < out.size:
            want = (out.size - filled) // 2
            draw = int(want * 1.35) + 16                       # acceptance probability pi / 4
            x = 2.0 * self._open_unit_interval(draw) - 1.0
            y = 2.0 * self._open_unit_interval(draw) - 1.0
            s = x * x + y * y
            ok = (s < 1.0) & (s > 0.0)
            x, y, s = x[ok][:want], y[ok][:want], s[ok][:want]
            f = numpy.sqrt(-2.0 * numpy.log(s) / s)
            out[filled:filled + x.size] = x * f
            out[filled + x.size:filled + 2 * x.size] = y * f
            filled += 2 * x.size
        return out[:count].reshape(self._shape(shape)) * std_dev


def _rand_gaussian_torus32(rng, message, sigma: float, shape, centered=False):
    rfloats = rng.gauss(shape, sigma)
    if centered:
        rfloats -= rfloats.mean()
    with numpy.errstate(over='ignore'):
        return (Torus32(message) + double_to_t32(rfloats)).astype(Torus32)


def rand_uniform_bool(thr, rng, shape):
    return thr.to_device(rng.uniform_bool(shape))


def rand_uniform_torus32(thr, rng, shape):
    return thr.to_device(rng.uniform_torus32(shape))


def rand_gaussian_torus32(thr, rng, message, sigma: float, shape, centered=False):
    return thr.to_device(_rand_gaussian_torus32(rng, message, sigma, shape, centered=centered))
