"""Batch sharding across the GPUs of one box: signals are independent,
every rank transforms a contiguous range of the batch with no collective on the data
path; the only communication is an optional final gather of the results.

Host-side plumbing only (torch.distributed); the transforms themselves are the C-ABI
calls of libfnft_b200.so."""


def shard_range(B, rank, world):
    """Contiguous, balanced partition of range(B): the first B % world ranks get one
    extra signal.  Returns (start, stop)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(B, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_rows(local_rows, B, group=None):
    """All-gathers per-rank result rows ([n_local, width] tensors, n_local given by
    shard_range) into a [B, width] tensor that is identical on every rank.  Works with
    the gloo (CPU tensors) and nccl (CUDA tensors) backends."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    start, stop = shard_range(B, rank, world)
    if local_rows.shape[0] != stop - start:
        raise ValueError("local_rows does not match this rank's shard")
    width = local_rows.shape[1]
    nmax = -(-B // world)
    # all_gather needs equal shapes: pad the short shards, trim after the exchange
    buf = torch.zeros((nmax, width), dtype=local_rows.dtype, device=local_rows.device)
    buf[:stop - start] = local_rows
    if local_rows.is_complex():
        parts = [torch.zeros((nmax, width, 2), dtype=local_rows.real.dtype, device=buf.device)
                 for _ in range(world)]
        dist.all_gather(parts, torch.view_as_real(buf).contiguous(), group=group)
        parts = [torch.view_as_complex(p) for p in parts]
    else:
        parts = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
    out = torch.empty((B, width), dtype=local_rows.dtype, device=local_rows.device)
    for r in range(world):
        s, e = shard_range(B, r, world)
        out[s:e] = parts[r][:e - s]
    return out
