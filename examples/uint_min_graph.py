"""Encrypted minimum of two arrays of 8-bit integers (the reference's uint_min circuit, operators_integer.py:64-95),
issued gate by gate and replayed as one CUDA graph (VirtualMachine.capture).  Prints both wall times.

    python examples/uint_min_graph.py [count]
"""
import os
import sys
import time

import numpy
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nufhe_b200 as nufhe                                                        # noqa: E402
from nufhe_b200.operators_integer import uint_min, uintarray_to_bitarray, bitarray_to_uintarray   # noqa: E402

count = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ctx = nufhe.Context(rng=nufhe.DeterministicRNG(11))
sk, ck = ctx.make_key_pair()
vm = ctx.make_virtual_machine(ck)
rng = numpy.random.RandomState(0)
xs, ys = rng.randint(0, 256, count).astype(numpy.uint8), rng.randint(0, 256, count).astype(numpy.uint8)
ca, cb = ctx.encrypt(sk, uintarray_to_bitarray(xs)), ctx.encrypt(sk, uintarray_to_bitarray(ys))
answer = vm.empty_ciphertext((count, 8))


def circuit():
    uint_min(ctx.thread, ck, answer, ca, cb, perf_params=vm.perf_params)


def wall(fn, reps=5):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t)
    return min(ts)


circuit()
t_eager = wall(circuit)
graph = vm.capture(circuit, reserve_batch=2 * count)
t_graph = wall(graph.replay)
got = bitarray_to_uintarray(ctx.decrypt(sk, answer))
assert (got == numpy.minimum(xs, ys)).all()
print('uint_min of %d x 8 bits (17 gates): eager %.2f ms, one graph launch %.2f ms' % (count, 1e3 * t_eager, 1e3 * t_graph))
