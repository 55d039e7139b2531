import torch


def arrays_equal(arr1, arr2):
    """nufhe/utils.py:17-20"""
    return arr1.shape == arr2.shape and bool(torch.equal(arr1.cpu(), arr2.cpu()))


def wrapping_dot(a, key):
    """(a * key).sum(-1) in Torus32 (wrap mod 2^32) -- vec_mul_mat, nufhe/lwe_cpu.py:22-23.
    a: (..., n) int32 device tensor, key: (n,) int32 device tensor."""
    s = (a.to(torch.int64) * key.to(torch.int64)).sum(-1)
    s = s & 0xffffffff
    s = torch.where(s >= 2**31, s - 2**32, s)
    return s.to(torch.int32)
