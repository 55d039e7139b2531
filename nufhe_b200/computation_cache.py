"""The reference caches JIT-compiled Reikna computations per Thread (nufhe/computation_cache.py).
This engine is compiled ahead of time, so there is nothing to cache; the entry point is kept."""


def clear_computation_cache(thr):
    return None
