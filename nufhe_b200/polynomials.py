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
