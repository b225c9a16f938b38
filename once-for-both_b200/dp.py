"""Data-parallel gradient exchange of the search step (reference: torch DDP wrap, search.py:617-620 - bucketed all-reduce
(mean) of every gradient once per optimizer step; survey 2.4).

The engine keeps every gradient (weights, biases, tokens, bi-mask scores and architecture alphas, decoder) in ONE flat fp32
arena, so the exchange is a handful of large NCCL all-reduces over NVLink/NVSwitch instead of DDP's per-parameter
bucketing: `bucket_bounds` cuts the arena into at most `max_buckets` contiguous, 16-byte aligned buckets of at least
`min_bucket_bytes` (launch latency, not link count, is what bucket size trades against on NVSwitch), and `allreduce_arena`
averages them in place. The functions are backend-agnostic (nccl on the GPUs, gloo in the CPU tests).
"""
from typing import List, Tuple

import torch
import torch.distributed as dist


def bucket_bounds(n_elems: int, max_buckets: int = 4, min_bucket_bytes: int = 8 << 20, elem_bytes: int = 4,
                  align_elems: int = 4) -> List[Tuple[int, int]]:
    """Contiguous [lo, hi) element ranges covering [0, n_elems)."""
    if n_elems <= 0:
        return []
    min_elems = max(align_elems, min_bucket_bytes // elem_bytes)
    nb = max(1, min(max_buckets, n_elems // min_elems))
    per = (n_elems + nb - 1) // nb
    per = (per + align_elems - 1) // align_elems * align_elems
    out, lo = [], 0
    while lo < n_elems:
        hi = min(n_elems, lo + per)
        out.append((lo, hi))
        lo = hi
    return out


def allreduce_arena(grads: torch.Tensor, world: int, group=None, bounds=None, async_op: bool = False):
    """In-place mean over ranks of the flat gradient arena. Returns the list of work handles when async_op."""
    if world <= 1:
        return []
    bounds = bounds if bounds is not None else bucket_bounds(grads.numel())
    backend = dist.get_backend(group)
    works = []
    for lo, hi in bounds:
        chunk = grads[lo:hi]
        if backend == "nccl":
            works.append(dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=group, async_op=async_op))
        else:   # gloo has no AVG
            w = dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
            works.append(w)
    if backend != "nccl":
        if async_op:
            for w in works:
                w.wait()
        grads.mul_(1.0 / world)
    return works if async_op else []
