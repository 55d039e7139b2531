"""C0 of BASELINE.md section 3: the reference's own CPU path -- the NumPy closures of nufhe/*_cpu.py composed as
nufhe/bootstrap.py prescribes -- timed on this machine's host cores.  It needs /root/reference, so it runs in the build
container only; the result is committed as profiles/r2_c0_reference_closures.json and quoted by bench.py next to the
C port (`cpu_baseline`).

    python tools/c0_baseline.py [ciphertexts] [steps]

The closures work on Python-object arrays (GaloisNumber) and hold the GIL: one core per process.  A bounded sample
(default 4 ciphertexts x 12 of the 500 blind-rotate steps, plus one full key switch) is timed and scaled to a whole
gate; the full 500-step run of the same closures is what tests/golden/make_golden*.py do (45 s per ciphertext)."""
import json
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import make_golden as M            # noqa: E402  (imports the unmodified reference through ref_bridge)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 12
N, n = 1024, 500
rng = numpy.random.RandomState(3)
bk_tr = rng.randint(0, 2**63, size=(STEPS, 2, 2, 2, N), dtype=numpy.int64).astype(numpy.uint64)
ks = (rng.randint(-2**31, 2**31, size=(N, 8, 4, n), dtype=numpy.int32),
      rng.randint(-2**31, 2**31, size=(N, 8, 4), dtype=numpy.int32), numpy.zeros((N, 8, 4), numpy.float32))
acc = rng.randint(-2**31, 2**31, size=(B, 2, N), dtype=numpy.int32)
bara = rng.randint(0, 2 * N, size=(B, n), dtype=numpy.int32)
shift = M.ShiftTorusPolynomialReference(N, (B, 2), (B, n), powers_view=True, minus_one=True)
extmul = M.TGswTransformedExternalMulReference(M.tgsw_params, (B,), STEPS, None)
t0 = time.perf_counter()
for i in range(STEPS):
    tmp = numpy.empty_like(acc)
    with numpy.errstate(over='ignore'):
        shift(tmp, acc, bara, i)
        extmul(tmp, bk_tr, i)
        acc = acc + tmp
t_steps = time.perf_counter() - t0
ea = rng.randint(-2**31, 2**31, size=(B, N), dtype=numpy.int32)
eb = rng.randint(-2**31, 2**31, size=(B,), dtype=numpy.int32)
ra, rb, rcv = numpy.empty((B, n), numpy.int32), numpy.empty((B,), numpy.int32), numpy.empty((B,), numpy.float32)
t0 = time.perf_counter()
with numpy.errstate(over='ignore'):
    M.LweKeyswitchReference(None, N, n, 8, 2)(ra, rb, rcv, ks[0], ks[1], ks[2], ea, eb)
t_ks = time.perf_counter() - t0
per_step = t_steps / (B * STEPS)
gate_s = 500 * per_step + t_ks / B
out = {'available': True, 'kind': 'reference', 'what': 'nufhe/*_cpu.py closures (tgsw_cpu.py:82-106, polynomials_cpu.py:25-59, '
       'lwe_cpu.py:62-93), composed per bootstrap.py:96-229', 'host': 'build container (no GPU)', 'cores': 1,
       'cores_on_host': os.cpu_count(),
       'sample': '%d ciphertexts x %d blind-rotate steps + 1 key switch, scaled to 500 steps' % (B, STEPS),
       'seconds_per_ciphertext_step': per_step, 'keyswitch_seconds_per_ciphertext': t_ks / B,
       'seconds_per_gate_nand': gate_s, 'gates_per_s_per_core': 1.0 / gate_s, 'unit': 'gates/s'}
print(json.dumps(out, indent=1))
with open(os.path.join(ROOT, 'profiles', 'r2_c0_reference_closures.json'), 'w') as f:
    json.dump(out, f, indent=1)
