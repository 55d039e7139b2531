"""Polynomial array containers (reference: nufhe/polynomials.py:30-104)."""
import pickle

import torch

from .utils import arrays_equal


class IntPolynomialArray:
    def __init__(self, coeffs):
        self.coeffs = coeffs
        self.polynomial_degree = coeffs.shape[-1]
        self.shape = tuple(coeffs.shape[:-1])


class TorusPolynomialArray:
    def __init__(self, coeffs):
        self.coeffs = coeffs
        self.polynomial_degree = coeffs.shape[-1]
        self.shape = tuple(coeffs.shape[:-1])

    @classmethod
    def empty(cls, thr, polynomial_degree, shape):
        return cls(thr.empty(tuple(shape) + (polynomial_degree,), torch.int32))


class TransformedPolynomialArray:
    """Transformed (NTT) polynomials: uint64 field elements stored as int64 bit patterns,
    natural order, length N (polynomial_transform_ntt.py:29-42)."""

    def __init__(self, transform_type, polynomial_degree, coeffs):
        self.transform_type = transform_type
        self.coeffs = coeffs
        self.polynomial_degree = polynomial_degree
        self.shape = tuple(coeffs.shape[:-1])

    @classmethod
    def empty(cls, thr, transform_type, polynomial_degree, shape):
        return cls(transform_type, polynomial_degree,
                   thr.empty(tuple(shape) + (polynomial_degree,), torch.int64))

    def dump(self, file_obj):
        pickle.dump(self.transform_type, file_obj)
        pickle.dump(self.polynomial_degree, file_obj)
        pickle.dump(self.coeffs.cpu().numpy().view('uint64'), file_obj)

    @classmethod
    def load(cls, file_obj, thr):
        transform_type = pickle.load(file_obj)
        polynomial_degree = pickle.load(file_obj)
        coeffs = pickle.load(file_obj)
        return cls(transform_type, polynomial_degree, thr.to_device(coeffs))

    def __eq__(self, other):
        return (self.__class__ == other.__class__ and self.transform_type == other.transform_type
                and self.polynomial_degree == other.polynomial_degree
                and arrays_equal(self.coeffs, other.coeffs))


def shift_tp_inverted_power(thr, result: TorusPolynomialArray, powers, source: TorusPolynomialArray):
    """result = X^(2N - pwr) * source, one power per polynomial (polynomials.py:90-95, K7 invert_powers)."""
    thr.shift_torus_polynomial(result.coeffs, source.coeffs, powers, mode=thr.SHIFT_INVERT)


def shift_tp_minus_one_power_from_array(thr, result: TorusPolynomialArray, powers, power_idx: int,
                                        source: TorusPolynomialArray):
    """result = (X^pwr - 1) * source with pwr = powers[..., power_idx], shared by the polynomials of one sample
    (polynomials.py:98-104, K7 powers_view + minus_one)."""
    thr.shift_torus_polynomial(result.coeffs, source.coeffs, powers, power_idx=power_idx,
                               polys_per_power=result.coeffs.shape[-2], mode=thr.SHIFT_MINUS_ONE)
