"""Data-parallel gradient exchange of the search step (reference: torch DDP wrap, search.py:617-620 - bucketed all-reduce
(mean) of every gradient once per optimizer step; survey 2.4).

The engine keeps every gradient (weights, biases, tokens, bi-mask scores and architecture alphas, decoder) in ONE flat fp32
arena, so the exchange is a handful of large NCCL all-reduces over NVLink/NVSwitch instead of DDP's per-parameter
bucketing: `bucket_bounds` cuts the arena into at most `max_buckets` contiguous, 16-byte aligned buckets of at least
`min_bucket_bytes` (launch latency, not link count, is what bucket size trades against on NVSwitch), and `allreduce_arena`
averages them in place. The engines exchange the arena in ONE piece after backward (8 x B200: 0.30 ms against 0.46 ms in four
buckets, profiles/r02b_allreduce_8gpu.txt); buckets only pay when the exchange is overlapped with backward. The functions are backend-agnostic (nccl on the GPUs, gloo in the CPU tests).
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


# ---------------------------------------------------------------------------------------------------------------------
# exchange overlapped with backward (north_star "bucketed gradient all-reduce ... overlapped with backward"; the reference
# gets the same from DDP's reducer hooks, search.py:619)
# ---------------------------------------------------------------------------------------------------------------------
def overlap_plan(n_elems: int, block_ranges: List[Tuple[int, int]], blocks_per_bucket: int = 2, tail_blocks: int = 2):
    """Bucket schedule of the flat gradient arena for a backward pass that finishes block depth-1 first and block 0 last.

    block_ranges[l] = [lo, hi) of the contiguous run holding the large (decay-group) weight gradients of block l; the runs
    ascend with l and are adjacent. Returns (early, tail):
      early  [(ready_after_block, lo, hi), ...] in launch order: the bucket may start as soon as the backward of block
             `ready_after_block` has been enqueued (all blocks >= it are complete);
      tail   [(lo, hi), ...] everything else - the small no-decay group, embedding / head / decoder tensors, the first
             `tail_blocks` blocks and the architecture parameters, whose gradients only complete at the very end.
    early + tail cover [0, n_elems) exactly once."""
    depth = len(block_ranges)
    early = []
    hi_blk = depth
    while hi_blk > tail_blocks:
        lo_blk = max(tail_blocks, hi_blk - blocks_per_bucket)
        early.append((lo_blk, block_ranges[lo_blk][0], block_ranges[hi_blk - 1][1]))
        hi_blk = lo_blk
    if not early:
        return [], [(0, n_elems)]
    first_lo = early[-1][1]
    last_hi = early[0][2]
    tail = [(lo, hi) for lo, hi in ((0, first_lo), (last_hi, n_elems)) if hi > lo]
    return early, tail


class OverlappedReducer:
    """Launches the bucket all-reduces of `overlap_plan` asynchronously while backward is still producing the remaining
    gradients, and joins them in `finish()`. Backend-agnostic: NCCL averages in the collective, gloo (CPU tests) sums and
    scales after the wait. One instance per engine; `begin()` at the start of every backward."""

    def __init__(self, grads: torch.Tensor, world: int, group, early, tail):
        self.grads, self.world, self.group = grads, world, group
        self.early, self.tail = list(early), list(tail)
        self.avg = world > 1 and dist.get_backend(group) == "nccl"
        self._works, self._launched = [], 0

    def begin(self):
        self._works, self._launched = [], 0

    def _launch(self, lo, hi):
        chunk = self.grads[lo:hi]
        op = dist.ReduceOp.AVG if self.avg else dist.ReduceOp.SUM
        self._works.append((dist.all_reduce(chunk, op=op, group=self.group, async_op=True), lo, hi))

    def on_block_done(self, l: int):
        """Backward of block l has been enqueued: launch every bucket that became complete."""
        if self.world <= 1:
            return
        while self._launched < len(self.early) and self.early[self._launched][0] >= l:
            _, lo, hi = self.early[self._launched]
            self._launch(lo, hi)
            self._launched += 1

    def finish(self):
        """All gradients are complete: launch what is left, wait for everything (stream-ordered on CUDA)."""
        if self.world <= 1:
            return
        while self._launched < len(self.early):
            _, lo, hi = self.early[self._launched]
            self._launch(lo, hi)
            self._launched += 1
        for lo, hi in self.tail:
            self._launch(lo, hi)
        for w, lo, hi in self._works:
            w.wait()
            if not self.avg:
                self.grads[lo:hi].mul_(1.0 / self.world)
        self._works = []
