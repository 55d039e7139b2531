"""The reference's examples/gate_nand.py on the B200 engine: same calls, `nufhe_b200` instead of `nufhe`."""
import os
import random
import sys
import time

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nufhe_b200 as nufhe   # noqa: E402

size = 4096
bits1 = [random.choice([False, True]) for i in range(size)]
bits2 = [random.choice([False, True]) for i in range(size)]
reference = [not (b1 and b2) for b1, b2 in zip(bits1, bits2)]

ctx = nufhe.Context()
secret_key, cloud_key = ctx.make_key_pair()

ciphertext1 = ctx.encrypt(secret_key, bits1)
ciphertext2 = ctx.encrypt(secret_key, bits2)

vm = ctx.make_virtual_machine(cloud_key)
result = vm.gate_nand(ciphertext1, ciphertext2)        # warm-up
ctx.thread.synchronize()
t = time.time()
result = vm.gate_nand(ciphertext1, ciphertext2)
ctx.thread.synchronize()
dt = time.time() - t
result_bits = ctx.decrypt(secret_key, result)

assert all(result_bits == reference)
print('%d NAND gates in %.1f ms (%.4f ms per bit)' % (size, dt * 1e3, dt * 1e3 / size))
