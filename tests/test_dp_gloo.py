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
