"""CUDA-graph capture of gate circuits.

A circuit such as `uint_min` (operators_integer.py:64-95) is a fixed sequence of small dependent gates -- per bit one
XNOR and one MUX on a (count, 1) slice -- i.e. dozens of kernel launches, fills, strided copies and temporary
allocations issued one by one from Python.  `GateGraph` records that sequence once into a `torch.cuda.CUDAGraph` (the
native launches go to torch's capturing stream, engine.py: Engine._call) and replays it with a single launch; inputs
and outputs are the ciphertext objects the circuit was captured with (refill them in place, then `replay()`).

The reference has no counterpart (every gate is a separate Reikna call); results are bit-identical to running the
circuit eagerly (tests/test_gpu_api.py)."""
import torch


class GateGraph:

    def __init__(self, thread, circuit, reserve_batch=None, warmup=1):
        """`circuit()` issues the gates (no host <-> device copies inside: upload operands before, read results after).
        It runs `warmup` times eagerly first -- that sizes the engine's scratch buffers and the allocator's pools, which
        must not grow during capture -- and then once more under capture."""
        self.thread = thread
        if reserve_batch:
            thread.reserve(reserve_batch)
        side = torch.cuda.Stream(device=thread.device)
        side.wait_stream(torch.cuda.current_stream(thread.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                circuit()
        torch.cuda.current_stream(thread.device).wait_stream(side)
        torch.cuda.synchronize(thread.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = circuit()

    def replay(self):
        self.graph.replay()
        return self.result
