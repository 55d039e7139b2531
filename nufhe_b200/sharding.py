"""Multi-GPU plumbing: ciphertext batches shard on the leading axis; the cloud key is broadcast once.

The reference's only multi-GPU model is "one Context per device, keys and ciphertext slices moved as
pickled bytes through the host" (examples/multi_gpu.py:46-104).  Here one process drives one GPU
(`torch.distributed`, NCCL over NVLink on the box, gloo in the CPU tests); there is no per-gate
collective because ciphertexts are independent."""
import torch
import torch.distributed as dist


def shard_bounds(batch, world, rank):
    """Contiguous, balanced split of `batch` items: ranks < batch % world get one extra item."""
    base, extra = divmod(batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def cloud_key_tensors(cloud_key):
    """The device arrays that make up a NuFHECloudKey, in a fixed order."""
    ks = cloud_key.keyswitch_key.lwe
    return [cloud_key.bootstrap_key.tgsw.samples.a.coeffs, ks.a, ks.b, ks.current_variances]


def broadcast_tensors(tensors, src=0, group=None):
    """Broadcast each tensor from `src` in place (one collective per tensor, set-up time only)."""
    for t in tensors:
        dist.broadcast(t, src=src, group=group)


def gather_shards(local, batch, world, rank, group=None):
    """All-gather variable-size leading-axis shards back into the full batch (host-side convenience for
    examples and tests; the benchmark never gathers inside the timed region)."""
    sizes = [shard_bounds(batch, world, r) for r in range(world)]
    maxlen = max(e - s for s, e in sizes)
    pad = torch.zeros((maxlen,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:e - s] for o, (s, e) in zip(out, sizes)], dim=0)
