"""Bootstrap key and the bootstrap procedure (reference: nufhe/bootstrap.py)."""
import pickle

import torch

from .numeric_functions import Torus32
from .lwe import LweParams, LweKey, LweSampleArray, LweKeyswitchKey, _keyswitch_into
from .tgsw import (
    TGswKey, TransformedTGswSampleArray, TGswParams, TGswSampleArray,
    tgsw_transform_samples, tgsw_encrypt_int, engine_format)


class BootstrapKey:
    """bootstrap.py:44-92"""

    def __init__(self, in_out_params: LweParams, tgsw: TransformedTGswSampleArray):
        bk_params = tgsw.params
        accum_params = bk_params.tlwe_params
        self.in_out_params = in_out_params
        self.bk_params = bk_params
        self.accum_params = accum_params
        self.extract_params = accum_params.extracted_lweparams
        self.tgsw = tgsw

    @classmethod
    def from_rng(cls, thr, rng, lwe_key: LweKey, tgsw_key: TGswKey, perf_params=None):
        in_out_params = lwe_key.params
        bk_params = tgsw_key.params
        accum_params = bk_params.tlwe_params
        bk = TGswSampleArray.empty(thr, bk_params, (in_out_params.size,))
        tgsw_encrypt_int(thr, rng, bk, lwe_key.key, accum_params.min_noise, tgsw_key, perf_params)
        bk_transformed = TransformedTGswSampleArray.empty(thr, bk_params, (in_out_params.size,))
        tgsw_transform_samples(thr, bk_transformed, bk, perf_params)
        return cls(in_out_params, bk_transformed)

    def dump(self, file_obj):
        pickle.dump(self.in_out_params, file_obj)
        self.tgsw.dump(file_obj)

    @classmethod
    def load(cls, file_obj, thr):
        in_out_params = pickle.load(file_obj)
        tgsw = TransformedTGswSampleArray.load(file_obj, thr)
        return cls(in_out_params, tgsw)

    def __eq__(self, other):
        return (self.__class__ == other.__class__ and self.in_out_params == other.in_out_params
                and self.tgsw == other.tgsw)


def _flat(sample: LweSampleArray):
    return (sample.a.contiguous(), sample.b.contiguous())


def bootstrap_affine(thr, result: LweSampleArray, bk: BootstrapKey, ks: LweKeyswitchKey, mu,
                     x1: LweSampleArray, x2, c, s1, s2, no_keyswitch=False):
    """bootstrap(mu, (0,c) + s1*x1 + s2*x2): the gates' linear prologue (gates.py:108-115), the
    mod-switch, test-vector rotation, 500-step blind rotation and sample extraction run as ONE kernel
    (nb_bootstrap_extract), followed by the key-switch kernel unless `no_keyswitch`."""
    bk_int = engine_format(thr, bk.tgsw)
    shape = tuple(result.shape)

    def expand(x):
        if x is None:
            return None
        a, b = x.a, x.b
        if tuple(b.shape) != shape:
            while b.dim() < len(shape):
                a, b = a.unsqueeze(0), b.unsqueeze(0)
            a, b = a.expand(shape + (a.shape[-1],)), b.expand(shape)
        return (a.contiguous(), b.contiguous())

    p1, p2 = expand(x1), expand(x2)
    if no_keyswitch:
        dense = result.a.is_contiguous() and result.b.is_contiguous()
        out = (result.a, result.b) if dense else None
        ext = thr.bootstrap_extract(p1, p2, c, s1, s2, mu, bk_int, out=out)
        if not dense:
            result.a.copy_(ext[0].reshape(result.a.shape))
            result.b.copy_(ext[1].reshape(result.b.shape))
        result.current_variances.zero_()
    else:
        ext = thr.bootstrap_extract(p1, p2, c, s1, s2, mu, bk_int)
        ext_sample = LweSampleArray(
            bk.extract_params, ext[0].reshape(shape + (ext[0].shape[-1],)), ext[1].reshape(shape),
            torch.zeros(shape, dtype=torch.float32, device=ext[1].device))
        _keyswitch_into(thr, result, ks, ext_sample, None, 0)


def bootstrap(thr, result: LweSampleArray, bk: BootstrapKey, ks: LweKeyswitchKey, mu,
              x: LweSampleArray, perf_params=None, no_keyswitch=False):
    """bootstrap.py:206-229: result = LWE(mu) iff phase(x) > 0, LWE(-mu) otherwise."""
    bootstrap_affine(thr, result, bk, ks, mu, x, None, 0, 1, 0, no_keyswitch=no_keyswitch)
