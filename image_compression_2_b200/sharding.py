"""Multi-GPU: the batch shards by image with NO collective on the hot path.

Each stream (one image's latent, fresh model) is independent, so rank r of G simply takes a
contiguous slice of the batch.  The only communication is the OPTIONAL gather of per-rank
compressed sizes, used to place every rank's streams in one global container; it is a few bytes
and latency-only, so it is a plain torch.distributed all_gather (NCCL on GPUs, gloo in CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(batch, rank, world_size):
    """Contiguous [lo, hi) slice of `batch` items for `rank`; remainders go to the low ranks."""
    base, rem = divmod(int(batch), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_shard_bytes(local_total_bytes, device=None):
    """All ranks learn every rank's compressed byte total.  Returns (int64 tensor [G] on `device`,
    this rank's global byte offset).  Works without an initialised process group (G = 1)."""
    t = torch.as_tensor([int(local_total_bytes)], dtype=torch.int64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t, 0
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    sizes = torch.cat(out)
    return sizes, int(sizes[: dist.get_rank()].sum())


def global_stream_offsets(local_offsets, rank_byte_offset):
    """Turn a rank's local stream offsets into offsets inside the global container."""
    return local_offsets + int(rank_byte_offset)
