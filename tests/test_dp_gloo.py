"""Host-side logic of the data-parallel gradient exchange on CPU (gloo, world_size 2): bucket cover, in-place mean,
and that two ranks with different gradients end up with identical, averaged arenas (what DDP guarantees, search.py:619)."""
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


def _worker(rank, world, port, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ofb_b200  # noqa: F401
    from ofb_b200 import dp
    g = torch.Generator().manual_seed(100 + rank)
    grads = torch.randn(n, generator=g)
    mine = grads.clone()
    bounds = dp.bucket_bounds(n, max_buckets=3, min_bucket_bytes=1024)
    dp.allreduce_arena(grads, world, None, bounds)
    torch.save({"mine": mine, "avg": grads}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_bucket_bounds_cover():
    import ofb_b200  # noqa: F401
    from ofb_b200 import dp
    for n in (4, 1000, 22_400_000, 5_900_004):
        b = dp.bucket_bounds(n)
        assert b[0][0] == 0 and b[-1][1] == n and len(b) <= 4
        assert all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
        assert all(lo % 4 == 0 for lo, _ in b)
    assert dp.bucket_bounds(0) == []
    assert len(dp.bucket_bounds(22_400_000)) == 4          # DeiT-S: 89.6 MB -> 4 buckets of 22.4 MB
    assert len(dp.bucket_bounds(1000)) == 1


def test_allreduce_arena_world2(tmp_path):
    world, n = 2, 10_007 * 4
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (torch.load(os.path.join(str(tmp_path), f"r{r}.pt")) for r in range(world))
    want = (r0["mine"] + r1["mine"]) / 2
    assert torch.allclose(r0["avg"], want, atol=1e-7) and torch.equal(r0["avg"], r1["avg"])


def test_single_rank_is_noop():
    import ofb_b200  # noqa: F401
    from ofb_b200 import dp
    g = torch.arange(16.0)
    assert dp.allreduce_arena(g, 1) == [] and torch.equal(g, torch.arange(16.0))


def _fake_block_ranges(depth, per_block, first):
    return [(first + l * per_block, first + (l + 1) * per_block) for l in range(depth)]


def test_overlap_plan_covers_arena_once():
    import ofb_b200  # noqa: F401
    from ofb_b200 import dp
    for depth, bpb, tailb in ((12, 2, 2), (12, 4, 0), (3, 2, 2), (2, 2, 2), (12, 5, 1)):
        n, first, per = 10_000, 1_000, 600
        early, tail = dp.overlap_plan(n, _fake_block_ranges(depth, per, first), bpb, tailb)
        cover = sorted([(lo, hi) for _, lo, hi in early] + tail)
        assert cover[0][0] == 0 and cover[-1][1] == n
        assert all(cover[i][1] == cover[i + 1][0] for i in range(len(cover) - 1))
        # launch order follows backward: ready blocks descend, and a bucket only holds blocks >= its ready block
        ready = [r for r, _, _ in early]
        assert ready == sorted(ready, reverse=True)
        for r, lo, hi in early:
            assert lo == first + r * per and r >= min(tailb, depth)
    # the real arena: block runs are adjacent and the plan covers it (engine builds it without touching the GPU kernels)
    early, tail = dp.overlap_plan(100, [], 2, 2)
    assert early == [] and tail == [(0, 100)]


def _overlap_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ofb_b200  # noqa: F401
    from ofb_b200 import dp
    depth, per, first, n = 6, 512, 256, 256 + 6 * 512 + 300
    g = torch.Generator().manual_seed(7 + rank)
    final = torch.randn(n, generator=g)            # what each region of the arena holds once its producer has run
    grads = torch.zeros(n)
    rng = _fake_block_ranges(depth, per, first)
    early, tail = dp.overlap_plan(n, rng, 2, 2)
    red = dp.OverlappedReducer(grads, world, None, early, tail)
    red.begin()
    for l in reversed(range(depth)):               # "backward": block l's gradients appear, then the hook fires
        lo, hi = rng[l]
        grads[lo:hi] = final[lo:hi]
        if l == 0:                                 # the small tensors complete last
            grads[:first] = final[:first]
            grads[rng[-1][1]:] = final[rng[-1][1]:]
        red.on_block_done(l)
    red.finish()
    torch.save({"mine": final, "avg": grads.clone()}, os.path.join(out_dir, f"o{rank}.pt"))
    dist.destroy_process_group()


def test_overlapped_reducer_world2(tmp_path):
    world = 2
    mp.spawn(_overlap_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (torch.load(os.path.join(str(tmp_path), f"o{r}.pt")) for r in range(world))
    want = (r0["mine"] + r1["mine"]) / 2
    assert torch.allclose(r0["avg"], want, atol=1e-7) and torch.equal(r0["avg"], r1["avg"])
