"""N-GPU check of the gradient exchange captured INSIDE the step graph (OFB_DP_GRAPH=1: bucket all-reduces + AdamW as graph
nodes, no host launches between backward and the update) against the host-launched exchange after the replay: identical
parameters on every rank and between the two modes (the gradients are deterministic), and the step time of both.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dp_graph_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import ofb_b200  # noqa: F401
from ofb_b200.engine import SearchStepEngine

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
D, H, depth, B = 384, 6, 12, 256
g = torch.Generator(device="cpu").manual_seed(1 + rank)
img = torch.randn(B, 3, 224, 224, generator=g).to(dev)
lab = torch.randint(0, 1000, (B,), generator=g).to(dev)


def run(in_graph, steps=3, timed=20):
    os.environ["OFB_DP_GRAPH"] = "1" if in_graph else "0"
    eng = SearchStepEngine(D, H, depth, B, drop_path_rate=0.1, lr=1e-3, device=dev, process_group=dist.group.WORLD)
    eng.init_params(seed=0)
    eng.set_schedule(0.0)
    torch.manual_seed(5)
    torch.cuda.manual_seed(5)
    for _ in range(steps):
        eng.step_graphed(img, lab)
    torch.cuda.synchronize()
    p = eng.params.clone()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(timed):
        eng.step_graphed(img, lab)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / timed], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    eng.release_graphs()
    torch.cuda.synchronize()
    return p, float(t.item())


ok = True
res = {}
for mode in (False, True, False, True):
    p, ms = run(mode)
    allp = [torch.empty_like(p) for _ in range(world)]
    dist.all_gather(allp, p)
    eq = all(torch.equal(allp[0], t) for t in allp)
    ok &= eq
    if mode in res:
        ok &= torch.equal(res[mode][0], p)
    res.setdefault(mode, (p, ms))
    if rank == 0:
        print(f"exchange in graph={mode}: ranks identical={eq}  {ms:.3f} ms/step ({world * B / ms * 1e3:.0f} img/s)", flush=True)
same = torch.equal(res[False][0], res[True][0])
if rank == 0:
    print(f"parameters after 3 steps identical between the two modes: {same}", flush=True)
    print("DP_GRAPH_CHECK", "OK" if (ok and same) else "FAILED", flush=True)
os._exit(0 if (ok and same) else 1)
