import torch


def arrays_equal(arr1, arr2):
    """nufhe/utils.py:17-20"""
    return arr1.shape == arr2.shape and bool(torch.equal(arr1.cpu(), arr2.cpu()))
