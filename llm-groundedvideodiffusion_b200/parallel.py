"""Multi-GPU plumbing for the replica-parallel hot path: one process per GPU, independent work
items (prompt index, repeat) per rank, no collective inside the denoising step.  `torch.distributed` (NCCL on the
B200 box, gloo in the CPU tests) carries only the packed work table and the gathered results / timings -- the same
partitioning the reference achieves by hand with `--skip_first_prompts/--num_prompts` (generate.py:36-53,255-262).
"""
import torch
import torch.distributed as dist


def shard_items(n_items, rank, world):
    """Contiguous windows (keeps the reference's directory / seed numbering): rank r owns [lo, hi)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def seed_for(prompt_index, repeat, seed_offset=0):
    """generate.py:241-243,325-335."""
    return prompt_index + repeat * 6789 + seed_offset


def broadcast_work_table(table, device):
    """Rank 0 holds `table` (int64 [n_items, k]: prompt index, repeat, seed, ...); every rank receives a copy."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return table.to(device)
    shape = torch.tensor(list(table.shape) if dist.get_rank() == 0 else [0, 0], dtype=torch.int64, device=device)
    dist.broadcast(shape, src=0)
    out = table.to(device) if dist.get_rank() == 0 else torch.empty(tuple(shape.tolist()), dtype=torch.int64,
                                                                    device=device)
    dist.broadcast(out, src=0)
    return out


def gather_results(local, n_items_total):
    """All ranks contribute a [n_local, ...] tensor of per-item results (e.g. uint8 frames, timings); rank 0 gets the
    concatenation in item order (contiguous windows => rank order), other ranks get None."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank()
    counts = [shard_items(n_items_total, r, world) for r in range(world)]
    m = max(hi - lo for lo, hi in counts)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0)
    if rank != 0:
        return None
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, counts)], dim=0)


def max_over_ranks(value, device):
    """Timing reduction used by bench.py: every rank passes its device-timed milliseconds, all get the max."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
