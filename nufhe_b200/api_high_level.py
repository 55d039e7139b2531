"""High-level API (reference: nufhe/api_high_level.py): find_devices, DeviceID, Context, VirtualMachine."""
import torch

from .api_low_level import (
    NuFHEParameters, NuFHESecretKey, NuFHECloudKey, encrypt, decrypt, empty_ciphertext)
from .lwe import LweSampleArray
from .random_numbers import DeterministicRNG
from .performance import PerformanceParameters
from .computation_cache import clear_computation_cache
from .gates import result_shape, get_shape
from .engine import Engine
from . import gates


def _check_api(api):
    if api not in (None, 'CUDA', 'cuda'):
        raise ValueError("Unrecognized API: " + str(api) + " (this engine is CUDA-only)")


def _match(name, include, exclude):
    if include is not None and not any(mask in name for mask in include):
        return False
    if exclude is not None and any(mask in name for mask in exclude):
        return False
    return True


def find_devices(api=None, include_devices=None, exclude_devices=None,
                 include_platforms=None, exclude_platforms=None):
    """api_high_level.py:45-82: the CUDA devices visible to this process that pass the name filters."""
    _check_api(api)
    if not _match('NVIDIA CUDA', include_platforms, exclude_platforms):
        return []
    ids = []
    for ordinal in range(torch.cuda.device_count()):
        name = torch.cuda.get_device_name(ordinal)
        if _match(name, include_devices, exclude_devices):
            ids.append(DeviceID('CUDA', 0, ordinal, 'NVIDIA CUDA', name))
    return ids


class DeviceID:
    """A picklable device identifier (api_high_level.py:85-127)."""

    def __init__(self, api_id, platform_id, device_id, platform_name='NVIDIA CUDA', device_name=''):
        self.api_id = api_id
        self.platform_id = platform_id
        self.device_id = device_id
        self.api_name = 'CUDA'
        self.platform_name = platform_name
        self.device_name = device_name

    def get_api_and_device(self):
        return 'CUDA', self.device_id

    def __str__(self):
        return "DeviceID(api={api}, platform={pnum} ({pname}), device={dnum} ({dname}))".format(
            api=self.api_name, pnum=self.platform_id, pname=self.platform_name,
            dnum=self.device_id, dname=self.device_name)


class Context:
    """An execution environment on one GPU (api_high_level.py:130-299).

    `thread` may be a ready nufhe_b200.engine.Engine; otherwise `device_id` (a DeviceID or a CUDA
    ordinal) selects the GPU, else the first device passing the filters.  `interactive` is accepted
    and ignored."""

    def __init__(self, rng=None, thread=None, device_id=None, api=None, interactive=False,
                 include_devices=None, exclude_devices=None,
                 include_platforms=None, exclude_platforms=None):
        if rng is None:
            rng = DeterministicRNG()
        if thread is not None:
            pass
        elif device_id is not None:
            ordinal = device_id.device_id if isinstance(device_id, DeviceID) else int(device_id)
            thread = Engine(ordinal)
        else:
            _check_api(api)
            if not torch.cuda.is_available():
                raise RuntimeError("nufhe_b200 needs a CUDA device; there is no CPU fallback")
            devices = find_devices(
                api=api, include_devices=include_devices, exclude_devices=exclude_devices,
                include_platforms=include_platforms, exclude_platforms=exclude_platforms)
            if len(devices) == 0:
                raise ValueError("No devices satisfying the given filters were found")
            thread = Engine(devices[0].device_id)
        self.rng = rng
        self.thread = thread

    def __del__(self):
        if hasattr(self, 'thread'):
            clear_computation_cache(self.thread)

    def make_secret_key(self, **params):
        return NuFHESecretKey.from_rng(self.thread, NuFHEParameters(**params), self.rng)

    def make_cloud_key(self, secret_key: NuFHESecretKey):
        return NuFHECloudKey.from_rng(self.thread, secret_key.params, self.rng, secret_key)

    def make_key_pair(self, **params):
        secret_key = self.make_secret_key(**params)
        cloud_key = self.make_cloud_key(secret_key)
        return secret_key, cloud_key

    def encrypt(self, secret_key: NuFHESecretKey, message):
        return encrypt(self.thread, self.rng, secret_key, message)

    def decrypt(self, secret_key: NuFHESecretKey, ciphertext: LweSampleArray):
        return decrypt(self.thread, secret_key, ciphertext)

    def make_virtual_machine(self, cloud_key: NuFHECloudKey, perf_params: PerformanceParameters = None):
        return VirtualMachine(self.thread, cloud_key, perf_params=perf_params)

    def load_ciphertext(self, file_or_bytestring):
        if isinstance(file_or_bytestring, bytes):
            return LweSampleArray.loads(file_or_bytestring, self.thread)
        return LweSampleArray.load(file_or_bytestring, self.thread)

    def load_secret_key(self, file_or_bytestring):
        if isinstance(file_or_bytestring, bytes):
            return NuFHESecretKey.loads(file_or_bytestring, self.thread)
        return NuFHESecretKey.load(file_or_bytestring, self.thread)

    def load_cloud_key(self, file_or_bytestring):
        if isinstance(file_or_bytestring, bytes):
            return NuFHECloudKey.loads(file_or_bytestring, self.thread)
        return NuFHECloudKey.load(file_or_bytestring, self.thread)


class VirtualMachine:
    """Executes gates on ciphertexts with an encapsulated cloud key (api_high_level.py:302-363).

    .. method:: gate_<operator>(*args, dest: LweSampleArray=None)
    """

    def __init__(self, thread, cloud_key: NuFHECloudKey, perf_params: PerformanceParameters = None):
        if perf_params is None:
            perf_params = PerformanceParameters(cloud_key.params)
        perf_params = perf_params.for_device(thread.device_params)
        self.thread = thread
        self.params = cloud_key.params
        self.cloud_key = cloud_key
        self.perf_params = perf_params
        # lay the bootstrap key out for the fused kernel once, at VM creation rather than at the first gate
        from .tgsw import engine_format, fused_kernel_supported
        if perf_params.single_kernel_bootstrap and fused_kernel_supported(cloud_key.bootstrap_key.bk_params):
            engine_format(thread, cloud_key.bootstrap_key.tgsw)

    def empty_ciphertext(self, shape):
        return empty_ciphertext(self.thread, self.params, shape)

    def load_ciphertext(self, file):
        return LweSampleArray.load(file, self.thread)

    def _gate(self, name, *args, dest: LweSampleArray = None):
        if dest is None:
            shapes = [get_shape(arg) for arg in args]
            dest = self.empty_ciphertext(result_shape(*shapes))
        gate_func = getattr(gates, name)
        gate_func(self.thread, self.cloud_key, dest, *args, perf_params=self.perf_params)
        return dest

    def capture(self, circuit, reserve_batch=None):
        """Record `circuit()` -- any sequence of this VM's gates on device-resident ciphertexts -- into a CUDA graph;
        returns a `GateGraph` whose `replay()` re-runs the whole circuit with one launch (nufhe_b200/graph.py)."""
        from .graph import GateGraph
        return GateGraph(self.thread, circuit, reserve_batch=reserve_batch)

    def __getattr__(self, name):
        if name.startswith('gate_'):
            return lambda *args, **kwds: self._gate(name, *args, **kwds)
        raise AttributeError(name)
