"""All-reduce microbenchmark of the gradient arena (run under torchrun): time of the 89.6 MB fp32 exchange for several
bucket counts, isolated from the step."""
import os
import sys

import torch
import torch.distributed as dist

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 22_400_000
g = torch.randn(n, device="cuda")
for nb in (1, 2, 4, 8):
    per = (n + nb - 1) // nb
    chunks = [g[i * per:min(n, (i + 1) * per)] for i in range(nb)]
    for _ in range(5):
        for c in chunks:
            dist.all_reduce(c, op=dist.ReduceOp.AVG)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        for c in chunks:
            dist.all_reduce(c, op=dist.ReduceOp.AVG)
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        ms = e0.elapsed_time(e1) / 20
        print(f"{os.environ.get('TAG', '')} buckets={nb}: {ms:.3f} ms per exchange, algbw {n * 4 / ms / 1e6:.0f} GB/s", file=sys.stderr, flush=True)
os._exit(0)
