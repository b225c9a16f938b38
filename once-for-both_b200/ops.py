"""Thin Python wrappers over the C ABI (include/ofb_b200.h).  Tensors in, tensors out; no math happens here.

Every function launches on the current torch CUDA stream and never synchronises.
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import GemmArgs, check as _check, cur_stream, lib, ptr

EPI_STORE, EPI_FC1, EPI_FC2_DGRAD, EPI_WGRAD, EPI_PATCH, EPI_DECODER = range(6)

# launch accounting (bench.py): every wrapper below is exactly one kernel launch of this library unless noted
LAUNCHES = 0
# optional per-launch GEMM timing: set to a list -> gemm() appends (start_event, end_event, flops)
GEMM_TIMING = None
# optional per-launch LayerNorm timing (the HBM-bound kernel family): list of (start_event, end_event, algorithmic bytes)
LN_TIMING = None


# optional algorithmic-work log (bench.py roofline): set to a list -> the GEMM / attention / LayerNorm wrappers append
# (kernel family, algorithmic FLOPs, algorithmic HBM bytes) per launch; durations come from CUPTI kernel records of graph replays
WORK_LOG = None


def _work(family, flops, nbytes, dims=None):
    if WORK_LOG is not None:
        WORK_LOG.append((family, float(flops), float(nbytes), dims))


def _ln_timed(call, nbytes):
    _work("ln", 0.0, nbytes)
    if LN_TIMING is None:
        return call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call()
    e1.record()
    LN_TIMING.append((e0, e1, nbytes))


# optional per-launch timing of EVERY library call (tools/step_breakdown.py): set to a list -> each launch appends
# (entry point, tag, start_event, end_event); the tag carries the GEMM shape / epilogue
OP_TIMING = None
_TAG = ""
_raw_lib = lib


class _TimedLib:
    def __init__(self, inner):
        self._inner = inner

    def __getattr__(self, name):
        fn = getattr(self._inner, name)
        if not name.startswith("ofb_") or name in ("ofb_version", "ofb_layernorm_bwd_parts", "ofb_gemm_mlp_partial_rows"):
            return fn

        def run(*args):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*args)
            e1.record()
            OP_TIMING.append((name, _TAG, e0, e1))
            return r
        return run


def lib():  # noqa: F811  (shadows the import on purpose: same object unless OP_TIMING is armed)
    return _TimedLib(_raw_lib()) if OP_TIMING is not None else _raw_lib()


# ---------------------------------------------------------------------------------------------------------------------
# deterministic split-K of the weight-gradient GEMMs (default on; OFB_DETERMINISTIC=0 restores red.global.add accumulation)
# ---------------------------------------------------------------------------------------------------------------------
DETERMINISTIC = os.environ.get("OFB_DETERMINISTIC", "1") == "1"


class _WgradPool:
    """Workspace of the split partial tiles + the pending fixed-order reductions. Inside `batch()` (the engines' backward) the
    reductions of several GEMMs are finished together by flush() - one launch per block; outside it every weight-gradient GEMM
    is finished right away.
    OFB_SPLITK_ASYNC=1 launches the reduction of a block on a SIDE stream (a parallel branch under CUDA-graph capture) so that
    it can run in the shadow of the next block's kernels: two workspace halves alternate, a half is reused only after the
    reduction that read it has finished (event), join() orders the caller's stream after every pending reduction. Measured on
    B200 it changes nothing (14.03 vs 14.02 ms per step: the persistent one-CTA-per-SM kernels leave the reduction no SM to
    hide on), so the default keeps the reductions on the caller's stream."""

    def __init__(self):
        self.bufs = {}          # device -> [half 0, half 1] fp32 tensors
        self.retired = []
        self.side = {}          # device -> side stream
        self.half = 0
        self.pending = [None, None]     # event of the reduction that last read each half
        self.cursor = 0
        self.batch_total = 0
        self.jobs = []
        self.batching = False
        self.async_reduce = os.environ.get("OFB_SPLITK_ASYNC", "0") == "1"

    def take(self, dev, nfloat):
        # batch_total counts everything taken since the last EXPLICIT flush: the halves are kept large enough for a whole batch,
        # so that a pass that was run eagerly once (growing them, possibly splitting a batch to do so) can be captured in a
        # CUDA graph afterwards without any allocation
        need = max(self.cursor, self.batch_total) + nfloat
        bufs = self.bufs.get(dev)
        if bufs is None or bufs[0].numel() < need:
            if torch.cuda.is_current_stream_capturing():
                raise _lib.OfbError("split-K workspace would have to grow during CUDA-graph capture: run one eager step first")
            if self.jobs:                       # pending partials live in the old buffers: finish them before they are replaced
                self._reduce()
            self.join()
            if bufs is not None:
                self.retired.append(bufs)       # CUDA graphs captured earlier still write their partials here: never freed
            n = max(need, 1 << 22) * 2
            bufs = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(2)]
            self.bufs[dev] = bufs
        if self.cursor == 0 and self.pending[self.half] is not None:
            torch.cuda.current_stream(dev).wait_event(self.pending[self.half])      # the reduction that read this half is done
            self.pending[self.half] = None
        off = self.cursor
        step = (nfloat + 3) // 4 * 4
        self.cursor += step
        self.batch_total += step
        return bufs[self.half][off:off + nfloat]

    def _launch(self, jobs):
        for i in range(0, len(jobs), 8):
            chunk = jobs[i:i + 8]
            arr = (_lib.SplitkJob * len(chunk))()
            for a, (ws, out, n4, splits, _dev) in zip(arr, chunk):
                a.ws, a.out, a.n4, a.splits = ws, out, n4, splits
            check(lib().ofb_splitk_reduce(C.cast(arr, C.c_void_p), len(chunk), cur_stream()), "ofb_splitk_reduce")

    def _reduce(self):
        jobs, self.jobs, self.cursor = self.jobs, [], 0
        if not jobs:
            return
        if not self.async_reduce:
            self._launch(jobs)
            return
        dev = jobs[0][4]
        cur = torch.cuda.current_stream(dev)
        side = self.side.get(dev)
        if side is None:
            side = self.side[dev] = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)                   # the partial tiles of this batch are complete
        with torch.cuda.stream(side):
            self._launch(jobs)
            ev = torch.cuda.Event()
            ev.record(side)
        self.pending[self.half] = ev
        self.half ^= 1

    def flush(self):
        self._reduce()
        self.batch_total = 0

    def join(self):
        """Order the current stream after every pending reduction (before anything reads the gradient arena)."""
        for i, ev in enumerate(self.pending):
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)
                self.pending[i] = None


_WG = _WgradPool()


class wgrad_batch:
    """with ops.wgrad_batch(): ... ops.wgrad_flush() ... - defer the split-K reductions to explicit flush points; on exit every
    reduction has been ordered before the caller's stream continues."""

    def __enter__(self):
        self.prev, _WG.batching = _WG.batching, True
        return self

    def __exit__(self, *exc):
        _WG.batching = self.prev
        if not _WG.batching and exc[0] is None:
            _WG.flush()
            _WG.join()
        return False


def wgrad_flush():
    _WG.flush()


def wgrad_join():
    _WG.join()


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


def check(code, what):
    _check(code, what)
    _count()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.OfbError("ofb_b200 kernels need CUDA tensors (no CPU fallback)")


def gemm(epi, A, B, *, M, N, K, out0=None, ld0=0, out1=None, ld1=0, out_fp32=False, bias=None, colscale=None,
         rowscale=None, rows_per_scale=1, res=None, aux=None, colpart0=None, colpart1=None, scale_ptr=None, pos=None,
         mask_token=None, rowmask=None, target=None, tokens=1, a_mn=False, b_mn=False, bn=0, k_splits=0, lda=None,
         ldb=None, bias_rowscaled=False, colscale_period=0):
    """D[M,N] = sum_k A[m,k] B[n,k] with a fused epilogue (see ofb_b200.h).  A/B are bf16, 2-D, last dim contiguous."""
    _need_cuda(A, B, out0)
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    g = GemmArgs()
    g.M, g.N, g.K, g.k_splits = M, N, K, k_splits
    g.out0, g.ld0 = ptr(out0), ld0 or (out0.stride(0) if out0 is not None and out0.dim() == 2 else 0)
    g.out1, g.ld1 = ptr(out1), ld1 or (out1.stride(0) if out1 is not None and out1.dim() == 2 else 0)
    g.out_fp32 = 1 if out_fp32 else 0
    g.bias_rowscaled = 1 if bias_rowscaled else 0
    g.bias, g.colscale, g.colscale_period = ptr(bias), ptr(colscale), colscale_period
    g.rowscale, g.rows_per_scale = ptr(rowscale), rows_per_scale
    g.res, g.ldres = ptr(res), (res.stride(0) if res is not None else 0)
    g.aux, g.ldaux = ptr(aux), (aux.stride(0) if aux is not None else 0)
    g.colpart0, g.colpart1, g.scale_ptr = ptr(colpart0), ptr(colpart1), ptr(scale_ptr)
    g.pos, g.mask_token, g.rowmask, g.target = ptr(pos), ptr(mask_token), ptr(rowmask), ptr(target)
    g.tokens = tokens
    lda = lda if lda is not None else A.stride(0)
    ldb = ldb if ldb is not None else B.stride(0)
    global _TAG
    _TAG = f"epi{epi} M{M} N{N} K{K} a{int(a_mn)}b{int(b_mn)}"
    _work("gemm", 2.0 * M * N * K, 2.0 * (M * K + N * K + M * N * (2 if epi == EPI_FC1 else 1)), (M, N, K, int(a_mn), int(b_mn), epi))
    det = False
    if epi == EPI_WGRAD and DETERMINISTIC and k_splits == 0 and N % 4 == 0 and g.ld0 == N:
        splits = _raw_lib().ofb_gemm_wgrad_splits(M, N, K, int(b_mn), bn)
        if splits > 1:          # a single split adds every element exactly once: already deterministic
            ws = _WG.take(out0.device, splits * M * N)
            g.k_splits, g.splitk_ws = splits, ptr(ws)
            _WG.jobs.append((ptr(ws), ptr(out0), M * N // 4, splits, out0.device))
            det = True
    if GEMM_TIMING is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib().ofb_gemm_bf16(epi, int(a_mn), int(b_mn), bn, ptr(A), lda, ptr(B), ldb, C.addressof(g), cur_stream()),
          "ofb_gemm_bf16")
    if GEMM_TIMING is not None:
        e1.record()
        GEMM_TIMING.append((e0, e1, 2.0 * M * N * K, epi))
    _TAG = ""
    if det and not _WG.batching:
        _WG.flush()
        _WG.join()


# ---------------------------------------------------------------------------------------------------------------------
# LayerNorm
# ---------------------------------------------------------------------------------------------------------------------
def layernorm_fwd(x, gamma, beta, y, mean, rstd, eps, d_valid=None):
    """d_valid: real channels of a zero-padded pruned embedding (None = all D)."""
    M, D = x.shape

    def call():
        if d_valid is None or d_valid == D:
            check(lib().ofb_layernorm_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd), M, D, eps, cur_stream()),
                  "ofb_layernorm_fwd")
        else:
            check(lib().ofb_layernorm_fwd_ex(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd), M, D, d_valid, eps,
                                             cur_stream()), "ofb_layernorm_fwd_ex")
    _ln_timed(call, 2.0 * M * D * 2)              # algorithmic bytes: read x, write y (bf16); SURVEY 8d


def gemm_mlp_partial_rows(n_tokens, bn):
    """Rows of the colpart buffers of the EPI_FC2_DGRAD epilogue (depends on how many epilogue column groups the kernel runs)."""
    r = _raw_lib().ofb_gemm_mlp_partial_rows(n_tokens, bn)
    if r <= 0:
        raise _lib.OfbError(f"ofb_gemm_mlp_partial_rows: invalid column tile {bn}")
    return r


def layernorm_bwd_parts(M):
    return lib().ofb_layernorm_bwd_parts(M)


def layernorm_bwd(dy, x, mean, rstd, gamma, dx, part_dgamma, part_dbeta, part_dbias=None, rowscale=None,
                  rows_per_scale=1, dres=None, d_valid=None):
    """dres: residual-branch gradient added to dx (pre-norm blocks); d_valid as in layernorm_fwd."""
    M, D = x.shape

    def call():
        if dres is None and (d_valid is None or d_valid == D):
            check(lib().ofb_layernorm_bwd(ptr(dy), ptr(x), ptr(mean), ptr(rstd), ptr(gamma), ptr(dx), ptr(part_dgamma),
                                          ptr(part_dbeta), ptr(part_dbias), ptr(rowscale), rows_per_scale, M, D, cur_stream()),
                  "ofb_layernorm_bwd")
        else:
            check(lib().ofb_layernorm_bwd_ex(ptr(dy), ptr(x), ptr(mean), ptr(rstd), ptr(gamma), ptr(dres), ptr(dx),
                                             ptr(part_dgamma), ptr(part_dbeta), ptr(part_dbias), ptr(rowscale), rows_per_scale,
                                             M, D, d_valid or D, cur_stream()), "ofb_layernorm_bwd_ex")
    _ln_timed(call, (3.0 + (1.0 if dres is not None else 0.0)) * M * D * 2)      # read dy, x (, dres), write dx


def reduce_partials(part, R, N, out, scale=1.0, div_by=None, accumulate=True):
    check(lib().ofb_reduce_partials(ptr(part), R, N, ptr(out), scale, ptr(div_by), int(accumulate), cur_stream()),
          "ofb_reduce_partials")


def reduce_partials_multi(jobs):
    """jobs: list of (part, R, N, out[, scale, div_by, accumulate]) tuples or dicts finished in ONE launch."""
    arr = (_lib.ReduceJob * len(jobs))()
    for a, j in zip(arr, jobs):
        if not isinstance(j, dict):
            j = dict(zip(("part", "R", "N", "out", "scale", "div_by", "accumulate"), j))
        a.part, a.out, a.div_by = ptr(j["part"]), ptr(j["out"]), ptr(j.get("div_by"))
        a.R, a.N, a.scale, a.accumulate = j["R"], j["N"], j.get("scale", 1.0), int(j.get("accumulate", True))
    check(lib().ofb_reduce_partials_multi(C.cast(arr, C.c_void_p), len(jobs), cur_stream()), "ofb_reduce_partials_multi")


# ---------------------------------------------------------------------------------------------------------------------
# token assembly / PMIM
# ---------------------------------------------------------------------------------------------------------------------
def patchify(images, patches, patch=16):
    B, _, HW, _ = images.shape
    check(lib().ofb_patchify(ptr(images), ptr(patches), B, HW, patch, cur_stream()), "ofb_patchify")


def _box(box):
    return (0, 0, 0, 0, 0) if box is None else (1,) + tuple(int(v) for v in box)


def mixup_batch(images, out, lam, box=None):
    """timm Mixup (batch mode) on the device: out <- lam x + (1-lam) x.flip(0), or the CutMix `box` = (yl, yh, xl, xh) pasted from
    x.flip(0). out may be `images` itself (in place, as timm does)."""
    _need_cuda(images, out)
    B, _, HW, _ = images.shape
    check(lib().ofb_mixup_batch(ptr(images), ptr(out), B, HW, float(lam), *_box(box), cur_stream()), "ofb_mixup_batch")


def patchify_mixup(images, patches, lam, box=None, patch=16):
    """patchify() of the mixed batch without materialising it."""
    _need_cuda(images, patches)
    B, _, HW, _ = images.shape
    check(lib().ofb_patchify_mixup(ptr(images), ptr(patches), B, HW, patch, float(lam), *_box(box), cur_stream()),
          "ofb_patchify_mixup")


def mixup_target(labels, target, lam, smoothing):
    _need_cuda(labels, target)
    B, Cn = target.shape
    check(lib().ofb_mixup_target(ptr(labels), ptr(target), B, Cn, float(lam), float(smoothing), cur_stream()), "ofb_mixup_target")


def pmim_mask(noise, mask, keep):
    B, L = noise.shape
    check(lib().ofb_pmim_mask(ptr(noise), ptr(mask), B, L, keep, cur_stream()), "ofb_pmim_mask")


def droppath_scale(u, drop_prob, scale):
    n_rows, B = u.shape
    check(lib().ofb_droppath_scale(ptr(u), ptr(drop_prob), ptr(scale), n_rows, B, cur_stream()), "ofb_droppath_scale")


def cls_rows(cls, pos, gate, x, B, T, D):
    check(lib().ofb_cls_rows(ptr(cls), ptr(pos), ptr(gate), ptr(x), B, T, D, cur_stream()), "ofb_cls_rows")


def embed_bwd(g0, x0, gate, mask, dconv, part_gx, part_pos, part_mt, B, T, D):
    check(lib().ofb_embed_bwd(ptr(g0), ptr(x0), ptr(gate), ptr(mask), ptr(dconv), ptr(part_gx), ptr(part_pos),
                              ptr(part_mt), B, T, D, cur_stream()), "ofb_embed_bwd")


def norm_targets(images, mask, target):
    B, _, HW, _ = images.shape
    check(lib().ofb_norm_targets(ptr(images), ptr(mask), ptr(target), B, HW, cur_stream()), "ofb_norm_targets")


# ---------------------------------------------------------------------------------------------------------------------
# losses / optimizer
# ---------------------------------------------------------------------------------------------------------------------
def ls_cross_entropy(logits, labels, loss_rows, dlogits, smoothing, grad_scale):
    B, Cn = logits.shape
    check(lib().ofb_ls_cross_entropy(ptr(logits), ptr(labels), ptr(loss_rows), ptr(dlogits), B, Cn, smoothing,
                                     grad_scale, cur_stream()), "ofb_ls_cross_entropy")


def soft_target_cross_entropy(logits, target, loss_rows, dlogits, grad_scale):
    B, Cn = logits.shape
    check(lib().ofb_soft_target_cross_entropy(ptr(logits), ptr(target), ptr(loss_rows), ptr(dlogits), B, Cn, grad_scale,
                                              cur_stream()), "ofb_soft_target_cross_entropy")


def eval_metrics(logits, labels, out_rows):
    B, Cn = logits.shape
    check(lib().ofb_eval_metrics(ptr(logits), ptr(labels), ptr(out_rows), B, Cn, cur_stream()), "ofb_eval_metrics")


def loss_finalize(loss_rows, dec_part, mask, arch_loss, grad_scale, scal):
    check(lib().ofb_loss_finalize(ptr(loss_rows), loss_rows.numel(), ptr(dec_part),
                                  dec_part.numel() if dec_part is not None else 0, ptr(mask),
                                  mask.numel() if mask is not None else 0, ptr(arch_loss), grad_scale, ptr(scal),
                                  cur_stream()), "ofb_loss_finalize")


def adamw(p, g, m, v, shadow, hyper, seg_end_host, zero_grad=True):
    """seg_end_host: ctypes int64 array of exclusive segment ends."""
    check(lib().ofb_adamw(ptr(p), ptr(g), ptr(m), ptr(v), ptr(shadow), ptr(hyper), len(seg_end_host),
                          C.cast(seg_end_host, C.c_void_p), p.numel(), int(zero_grad), cur_stream()), "ofb_adamw")


def colsum_bf16(x, R, N, out, scale=1.0, scale_dev=None, ld=None):
    check(lib().ofb_colsum_bf16(ptr(x), ld if ld is not None else x.stride(0), R, N, ptr(out), scale, ptr(scale_dev),
                                cur_stream()), "ofb_colsum_bf16")


class HyperUploader:
    """Per-step upload of the small hyper-parameter vector (lr, AdamW bias corrections, w_p) from a RING of pinned host slots by
    a kernel (ofb_copy_f32 reads the pinned slot directly). The ring keeps a host that runs several steps ahead of the GPU from
    overwriting values an enqueued upload has not read yet (each slot's upload is fenced by an event before the slot is
    refilled); the kernel keeps the upload off the copy engines, where it would queue behind the next batch's bulk H2D copy."""

    def __init__(self, n, device, slots=8):
        import torch
        self.n, self.dev = n, device
        cuda = torch.device(device).type == "cuda"
        self.slots = [torch.zeros(n, dtype=torch.float32).pin_memory() if cuda else torch.zeros(n) for _ in range(slots)]
        self.events = [None] * slots
        self.cur = 0

    @property
    def host(self):
        return self.slots[self.cur]

    def begin(self, keep=False):
        """Advance to the next slot (waiting, normally not at all, until its previous upload has been consumed) and return it;
        keep=True starts from the previous slot's values."""
        prev = self.slots[self.cur]
        self.cur = (self.cur + 1) % len(self.slots)
        if self.events[self.cur] is not None:
            self.events[self.cur].synchronize()
        if keep:
            self.slots[self.cur].copy_(prev)
        return self.slots[self.cur]

    def upload(self, dst):
        import torch
        _need_cuda(dst)
        check(lib().ofb_copy_f32(self.slots[self.cur].data_ptr(), ptr(dst), self.n, cur_stream()), "ofb_copy_f32")
        if self.events[self.cur] is None:
            self.events[self.cur] = torch.cuda.Event()
        self.events[self.cur].record()


def cast_bf16(src, dst):
    check(lib().ofb_cast_bf16(ptr(src), ptr(dst), src.numel(), cur_stream()), "ofb_cast_bf16")


# ---------------------------------------------------------------------------------------------------------------------
# bi-mask
# ---------------------------------------------------------------------------------------------------------------------
def bimask_fwd(mods_dev, nmod, max_n, params, switches, widths, w_p_dev, gate, rank, aprob, wsum, sp_loss):
    check(lib().ofb_bimask_fwd(ptr(mods_dev), nmod, max_n, ptr(params), ptr(switches), ptr(widths), ptr(w_p_dev),
                               ptr(gate), ptr(rank), ptr(aprob), ptr(wsum), ptr(sp_loss), cur_stream()), "ofb_bimask_fwd")


def arch_finalize(mods_dev, nmod, wsum, sp_loss, depth, D, H, d, hidden, L, Cn, target_flops, w_flops, arch, dwsum, d_active=0):
    """D, H, d, hidden: original dims; d_active: current LayerNorm width after truncating prune events (0 = D)."""
    check(lib().ofb_arch_finalize(ptr(mods_dev), nmod, ptr(wsum), ptr(sp_loss), depth, D, H, d, hidden, d_active, L, Cn,
                                  target_flops, w_flops, ptr(arch), ptr(dwsum), cur_stream()), "ofb_arch_finalize")


def bimask_bwd(mods_dev, nmod, max_n, params, switches, widths, w_p_dev, dgate, rank, aprob, dwsum, grad_scale, grads):
    check(lib().ofb_bimask_bwd(ptr(mods_dev), nmod, max_n, ptr(params), ptr(switches), ptr(widths), ptr(w_p_dev),
                               ptr(dgate), ptr(rank), ptr(aprob), ptr(dwsum), grad_scale, ptr(grads), cur_stream()),
          "ofb_bimask_bwd")


# ---------------------------------------------------------------------------------------------------------------------
# attention
# ---------------------------------------------------------------------------------------------------------------------
def attention_fwd(qkv, o, lse, drop_scale, B, T, H, scale):
    # algorithmic work: 4 H T^2 d FLOPs per image; HBM: read q, k, v, write o (bf16) + lse
    _work("attn_fwd", 4.0 * B * H * T * T * 64, 2.0 * B * T * H * 64 * 4 + 4.0 * B * H * T)
    check(lib().ofb_attention_fwd(ptr(qkv), ptr(o), ptr(lse), ptr(drop_scale), B, T, H, scale, cur_stream()),
          "ofb_attention_fwd")


def attention_bwd(qkv, o, d_o, lse, gate, drop_scale, dqkv, part_gate, part_bias, B, T, H, scale):
    # backward = 2.5 x forward FLOPs (S recomputed, dP, dV, dK, dQ); HBM: read q, k, v, o, dO, write dq, dk, dv
    _work("attn_bwd", 10.0 * B * H * T * T * 64, 2.0 * B * T * H * 64 * 8 + 4.0 * B * H * T)
    check(lib().ofb_attention_bwd(ptr(qkv), ptr(o), ptr(d_o), ptr(lse), ptr(gate), ptr(drop_scale), ptr(dqkv),
                                  ptr(part_gate), ptr(part_bias), B, T, H, scale, cur_stream()), "ofb_attention_bwd")
