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


def plan_work_items(n_prompts, save_dir=None, *, repeats=1, seed_offset=0, skip_first_prompts=0, num_prompts=None,
                    regenerate=1, seed_overrides=None):
    """The work list of one `generate.py` run (generate.py:225-356), as rows (ind, prompt index, repeat, seed):
      * `ind` numbers the output directory `<save_dir>/<ind>` and runs over regenerate x prompts, also counting the
        prompts skipped by `skip_first_prompts` / `num_prompts` (:254-262) -- so windows of a sharded run keep the
        directory and seed numbering of a single-process run;
      * seed = ind + repeat * 6789 + seed_offset, or seed_override + regenerate_ind * 56789 + ... when the prompt carries
        its own seed (:241-243,320-335);
      * RESUME: a directory that already holds `repeats` files `video*.joblib` is complete and skipped (:289-299); inside
        an incomplete directory the repeats whose `video_<repeat>.gif` exists are skipped (generation/lvd.py:150-153)."""
    import os
    rows = []
    ind = 0
    if regenerate > 1:
        assert skip_first_prompts == 0  # generate.py:246-248
    for regen in range(regenerate):
        for p in range(n_prompts):
            this = ind
            ind += 1
            if p < skip_first_prompts or (num_prompts is not None and p >= skip_first_prompts + num_prompts):
                continue
            d = os.path.join(save_dir, str(this)) if save_dir is not None else None
            if d is not None and os.path.isdir(d):
                done = [f for f in os.listdir(d) if f.startswith("video") and f.endswith("joblib")]
                if len(done) >= repeats:
                    continue
            override = seed_overrides.get(p) if seed_overrides else None
            base = override + regen * 56789 if override is not None else this
            for r in range(repeats):
                if d is not None and os.path.exists(os.path.join(d, f"video_{r}.gif")):
                    continue
                rows.append((this, p, r, base + r * 6789 + seed_offset))
    return rows


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


def generate_sharded(generate_fn, layouts, device, *, repeats=1, seed_offset=0, stats=None, batch=1):
    """The batch-of-prompts driver of SURVEY.md section 8(e): rank 0 holds `layouts` (one entry per prompt: a dict with
    `bboxes` [n_obj][F][4] floats and `object_positions` [n_obj][n_tok] ints); the work table (prompt index, repeat,
    seed) and the packed box tables are broadcast, every rank generates its contiguous window with
    `generate_fn(seed, bboxes, object_positions) -> uint8 [F, H, W, 3]` (the pipeline call; no collective inside), and
    the frames are gathered on rank 0 in item order.  Returns (frames uint8 [n_items, F, H, W, 3] | None, table).
    `stats` (a dict) receives this rank's wall-clock seconds of the three phases: scatter, generate, gather.
    batch > 1: the rank's window is cut into groups of up to `batch` items that go through the pipeline together --
    `generate_fn(seeds, bboxes_per_item, positions_per_item) -> uint8 [n, F, H, W, 3]` receives lists (multi-prompt
    batching: one UNet call per step for the whole group, guidance in lock-step)."""
    import time
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0

    def _sync():
        if torch.device(device).type == "cuda":
            torch.cuda.synchronize(device)
        return time.perf_counter()

    t_start = _sync()
    if rank == 0:
        n_prompts = len(layouts)
        rows = [(p, r, seed_for(p, r, seed_offset)) for p in range(n_prompts) for r in range(repeats)]
        table = torch.tensor(rows, dtype=torch.int64).reshape(-1, 3)
        max_obj = max(len(l["bboxes"]) for l in layouts)
        max_tok = max(len(t) for l in layouts for t in l["object_positions"])
        frames = len(layouts[0]["bboxes"][0])
        boxes = torch.zeros((n_prompts, max_obj, frames, 4), dtype=torch.float64)  # float64: box rounding stays bit-exact
        toks = torch.full((n_prompts, max_obj, max_tok), -1, dtype=torch.int64)
        for p, l in enumerate(layouts):
            for o, (ob, ot) in enumerate(zip(l["bboxes"], l["object_positions"])):
                boxes[p, o] = torch.tensor(ob, dtype=torch.float64)
                toks[p, o, :len(ot)] = torch.tensor(ot, dtype=torch.int64)
        meta = torch.tensor([n_prompts, max_obj, max_tok, frames], dtype=torch.int64)
    else:
        table = torch.empty(0, dtype=torch.int64)
        meta = torch.zeros(4, dtype=torch.int64)
        boxes = toks = None
    table = broadcast_work_table(table, device)
    if world > 1:
        meta = meta.to(device)
        dist.broadcast(meta, src=0)
        n_prompts, max_obj, max_tok, frames = (int(v) for v in meta.tolist())
        boxes = boxes.to(device) if rank == 0 else torch.empty((n_prompts, max_obj, frames, 4), dtype=torch.float64,
                                                               device=device)
        toks = toks.to(device) if rank == 0 else torch.empty((n_prompts, max_obj, max_tok), dtype=torch.int64,
                                                             device=device)
        dist.broadcast(boxes, src=0)
        dist.broadcast(toks, src=0)
    n_items = table.shape[0]
    lo, hi = shard_items(n_items, rank, world)
    t_scattered = _sync()
    outs = []
    items = []
    for i in range(lo, hi):
        p, _, seed = (int(v) for v in table[i].tolist())
        pos = [[int(t) for t in row.tolist() if t >= 0] for row in toks[p]]
        keep = [o for o, t in enumerate(pos) if len(t) > 0]
        bb = [[tuple(float(v) for v in fr) for fr in boxes[p, o].tolist()] for o in keep]
        items.append((seed, bb, [pos[o] for o in keep]))
    if batch <= 1:
        for seed, bb, pos in items:
            outs.append(torch.as_tensor(generate_fn(seed, bb, pos)))
    else:
        for k in range(0, len(items), batch):
            grp = items[k:k + batch]
            res = torch.as_tensor(generate_fn([g[0] for g in grp], [g[1] for g in grp], [g[2] for g in grp]))
            assert res.shape[0] == len(grp)
            outs.extend(res[j] for j in range(len(grp)))
    if outs:
        local = torch.stack(outs).to(device)
    else:  # a rank with an empty window still takes part in the gather (shape from a broadcast of rank 0's)
        local = None
    shape = torch.tensor(list(local.shape[1:]) if local is not None else [0, 0, 0, 0], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(shape, op=dist.ReduceOp.MAX)
    if local is None:
        local = torch.empty((0,) + tuple(int(v) for v in shape.tolist()), dtype=torch.uint8, device=device)
    t_generated = _sync()
    out = gather_results(local, n_items)
    t_done = _sync()
    if stats is not None:
        stats.update(scatter_s=t_scattered - t_start, generate_s=t_generated - t_scattered, gather_s=t_done - t_generated,
                     items=hi - lo)
    return out, table
