"""CPU test of the N>1 host path: world_size-2 gloo processes exercise the work-table broadcast, the contiguous
sharding, the ordered gather and the max-over-ranks timing reduction that bench.py uses under NCCL."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lvd_b200 import parallel
    n_items = 7
    table = torch.stack([torch.arange(n_items), torch.zeros(n_items, dtype=torch.int64),
                         torch.tensor([parallel.seed_for(i, 0, 3) for i in range(n_items)])], 1) \
        if rank == 0 else torch.empty(0, dtype=torch.int64)
    table = parallel.broadcast_work_table(table, "cpu")
    lo, hi = parallel.shard_items(n_items, rank, world)
    local = (table[lo:hi, 2:3].float() * 2.0)  # "result" of each item
    out = parallel.gather_results(local, n_items)
    mx = parallel.max_over_ranks(10.0 + rank, "cpu")
    if rank == 0:
        ret["gathered"] = out.flatten().tolist()
    ret[f"max{rank}"] = mx
    ret[f"win{rank}"] = (lo, hi)
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_gather():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret["win0"] == (0, 4) and ret["win1"] == (4, 7)
    assert ret["gathered"] == [2.0 * (i + 3) for i in range(7)]
    assert ret["max0"] == 11.0 and ret["max1"] == 11.0


def test_shard_items_covers_everything():
    from lvd_b200 import parallel
    for n in (0, 1, 7, 32, 500):
        for w in (1, 2, 4, 8):
            wins = [parallel.shard_items(n, r, w) for r in range(w)]
            assert wins[0][0] == 0 and wins[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(wins, wins[1:]))
            assert max(h - l for l, h in wins) - min(h - l for l, h in wins) <= 1


def _gen_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lvd_b200 import parallel
    F = 4
    layouts = None
    if rank == 0:
        layouts = [dict(bboxes=[[(0.1 * p, 0.2, 0.5, 0.6 + 0.01 * f) for f in range(F)]] * (1 + p % 2),
                        object_positions=[[2], [4, 5]][: 1 + p % 2]) for p in range(5)]
    seen = []

    def fake_pipe(seed, bboxes, positions):  # stands in for the pipeline call: encodes its inputs into the "frames"
        seen.append((seed, len(bboxes), [len(t) for t in positions]))
        img = torch.zeros((F, 2, 2, 3), dtype=torch.uint8)
        img[..., 0] = seed % 251
        img[..., 1] = len(bboxes)
        img[..., 2] = int(round(bboxes[0][0][0] * 100))
        return img

    frames, table = parallel.generate_sharded(fake_pipe, layouts, "cpu", repeats=2, seed_offset=1)
    ret[f"seen{rank}"] = seen
    if rank == 0:
        ret["frames"] = frames[:, 0, 0, 0, :].tolist()
        ret["table"] = table.tolist()
    dist.destroy_process_group()


def test_generate_sharded_two_ranks_gloo():
    """Work table + packed box tables broadcast from rank 0, contiguous windows, frames gathered in item order."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gen_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    table = ret["table"]
    assert len(table) == 10 and table[3] == [1, 1, 1 + 6789 + 1]
    assert len(ret["seen0"]) == 5 and len(ret["seen1"]) == 5
    for (p, r, seed), px in zip(table, ret["frames"]):
        assert px == [seed % 251, 1 + p % 2, 10 * p]
    assert ret["seen1"][0] == (table[5][2], 1 + table[5][0] % 2, [1, 2][: 1 + table[5][0] % 2])


def _gen_batch_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lvd_b200 import parallel
    F = 4
    layouts = None
    if rank == 0:
        layouts = [dict(bboxes=[[(0.1 * p, 0.2, 0.5, 0.6) for f in range(F)]] * (1 + p % 3),
                        object_positions=[[2], [4, 5], [7]][: 1 + p % 3]) for p in range(7)]
    groups = []

    def fake_pipe(seeds, bboxes, positions):  # multi-prompt form: lists in, [n, F, H, W, 3] out
        groups.append(list(seeds))
        out = torch.zeros((len(seeds), F, 2, 2, 3), dtype=torch.uint8)
        for j, (s, bb) in enumerate(zip(seeds, bboxes)):
            out[j, ..., 0] = s % 251
            out[j, ..., 1] = len(bb)
        return out

    stats = {}
    frames, table = parallel.generate_sharded(fake_pipe, layouts, "cpu", stats=stats, batch=3)
    ret[f"groups{rank}"] = groups
    ret[f"stats{rank}"] = dict(stats)
    if rank == 0:
        ret["frames"] = frames[:, 0, 0, 0, :2].tolist()
    dist.destroy_process_group()


def test_generate_sharded_multi_prompt_groups_gloo():
    """batch=3: each rank's window goes through the pipeline in groups of up to 3 items; order and content of the
    gathered frames are unchanged; phase timings are reported."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gen_batch_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret["groups0"] == [[0, 1, 2], [3]] and ret["groups1"] == [[4, 5, 6]]
    assert ret["frames"] == [[p % 251, 1 + p % 3] for p in range(7)]
    for r in (0, 1):
        st = ret[f"stats{r}"]
        assert set(st) == {"scatter_s", "generate_s", "gather_s", "items"} and st["items"] == (4 if r == 0 else 3)


def test_plan_work_items_seed_schedule_and_resume(tmp_path):
    """generate.py:225-356: directory / seed numbering survives windows, completed directories are skipped, repeats whose
    gif exists are skipped, per-prompt seed overrides follow the regenerate schedule."""
    from lvd_b200 import parallel
    rows = parallel.plan_work_items(5, None, repeats=2, seed_offset=3)
    assert rows[:3] == [(0, 0, 0, 3), (0, 0, 1, 6789 + 3), (1, 1, 0, 1 + 3)] and len(rows) == 10
    # a window keeps the global numbering
    win = parallel.plan_work_items(5, None, repeats=1, skip_first_prompts=2, num_prompts=2)
    assert win == [(2, 2, 0, 2), (3, 3, 0, 3)]
    # resume
    (tmp_path / "1").mkdir()
    (tmp_path / "1" / "video_0.joblib").write_bytes(b"x")
    (tmp_path / "1" / "video_1.joblib").write_bytes(b"x")
    (tmp_path / "3").mkdir()
    (tmp_path / "3" / "video_0.joblib").write_bytes(b"x")
    (tmp_path / "3" / "video_0.gif").write_bytes(b"x")
    rows = parallel.plan_work_items(5, str(tmp_path), repeats=2)
    assert [r[:3] for r in rows] == [(0, 0, 0), (0, 0, 1), (2, 2, 0), (2, 2, 1), (3, 3, 1), (4, 4, 0), (4, 4, 1)]
    # regenerate with a per-prompt seed override (prompt 1 carries seed 100)
    rows = parallel.plan_work_items(2, None, repeats=1, regenerate=2, seed_overrides={1: 100})
    assert rows == [(0, 0, 0, 0), (1, 1, 0, 100), (2, 0, 0, 2), (3, 1, 0, 100 + 56789)]
