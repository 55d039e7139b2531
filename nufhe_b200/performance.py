"""PerformanceParameters (reference: nufhe/performance.py:22-236).

The reference uses these to pick JIT variants (arithmetic flavour, constant memory, transforms per
block, single- vs multi-kernel bootstrap).  The B200 engine has one ahead-of-time compiled path, so
the object is accepted everywhere the reference accepts it, is value-hashable like the reference's,
and its knobs are validated but otherwise ignored."""


class PerformanceParameters:

    __attributes__ = (
        'nufhe_params', 'ntt_base_method', 'ntt_mul_method', 'ntt_lsh_method',
        'use_constant_memory_multi_iter', 'use_constant_memory_single_iter',
        'transforms_per_block', 'single_kernel_bootstrap', 'low_end_device')

    def __init__(
            self, nufhe_params, ntt_base_method=None, ntt_mul_method=None, ntt_lsh_method=None,
            use_constant_memory_multi_iter=None, use_constant_memory_single_iter=None,
            transforms_per_block=None, single_kernel_bootstrap=None, low_end_device=None):
        assert ntt_base_method in (None, 'cuda_asm', 'c')
        assert ntt_mul_method in (None, 'cuda_asm', 'c_from_asm', 'c')
        assert ntt_lsh_method in (None, 'cuda_asm', 'c_from_asm', 'c')
        self.nufhe_params = nufhe_params
        self.ntt_base_method = ntt_base_method
        self.ntt_mul_method = ntt_mul_method
        self.ntt_lsh_method = ntt_lsh_method
        self.use_constant_memory_multi_iter = use_constant_memory_multi_iter
        self.use_constant_memory_single_iter = use_constant_memory_single_iter
        self.transforms_per_block = transforms_per_block
        self.single_kernel_bootstrap = single_kernel_bootstrap
        self.low_end_device = low_end_device

    def for_device(self, device_params):
        return PerformanceParametersForDevice(self, device_params)

    def _key(self):
        return tuple(getattr(self, attr) for attr in self.__attributes__)

    def __hash__(self):
        return hash((self.__class__,) + self._key())

    def __eq__(self, other):
        return self.__class__ == other.__class__ and self._key() == other._key()


def single_kernel_bootstrap_supported(nufhe_params, device_params=None):
    """blind_rotate.py:37-86: the fused kernel exists for mask size 1 and the default decomposition only."""
    from .tgsw import fused_kernel_supported
    return fused_kernel_supported(nufhe_params.tgsw_params)


class PerformanceParametersForDevice:
    """performance.py:137-236.  `single_kernel_bootstrap` defaults to True (the fused kernel); False selects the
    reference's multi-kernel sequence of separate launches (bootstrap.py:96-196), same results, much slower."""

    def __init__(self, perf_params: PerformanceParameters, device_params):
        self.nufhe_params = perf_params.nufhe_params
        self.ntt_base_method = 'cuda_asm'
        self.ntt_mul_method = 'cuda_asm'
        self.ntt_lsh_method = 'cuda_asm'
        self.use_constant_memory_multi_iter = False
        self.use_constant_memory_single_iter = False
        self.transforms_per_block = 4
        supported = single_kernel_bootstrap_supported(perf_params.nufhe_params, device_params)
        if perf_params.single_kernel_bootstrap is None:
            self.single_kernel_bootstrap = supported
        else:
            if perf_params.single_kernel_bootstrap and not supported:           # performance.py:183-185
                raise ValueError("Single kernel bootstrap is not supported for this parameter set")
            self.single_kernel_bootstrap = bool(perf_params.single_kernel_bootstrap)
        self.low_end_device = False

    def _key(self):
        return (self.nufhe_params, self.single_kernel_bootstrap)

    def __hash__(self):
        return hash((self.__class__,) + self._key())

    def __eq__(self, other):
        return self.__class__ == other.__class__ and self._key() == other._key()
