"""TEST INFRASTRUCTURE ONLY — CPU oracle: a plain fp32 PyTorch restatement of the Once-for-Both bi-mask DeiT search
step (forward, losses, backward through autograd, AdamW).  It is the checker for the CUDA path; nothing in the
product (once-for-both_b200/) may import it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg use it.

Parity pin: oracle/make_golden.py runs the UNMODIFIED reference (/root/reference through oracle/ref_shim.py) and this
restatement on identical seeded parameters / inputs / noise, asserts agreement (logits, every loss term, every
gradient, post-AdamW parameters) and stores the reference's outputs in tests/golden/*.npz.  tests/test_oracle_golden.py
re-checks the restatement against those fixtures on every run (no /root/reference needed).

The restatement is functional (parameters in a dict keyed by the reference's state_dict names) and is written from the
closed forms in SURVEY.md App. A rather than from the reference's op sequence:
  * the softmax(alpha)-weighted prefix masks are suffix sums over the search-space cells,
  * rank lookups are done by counting instead of double argsort,
so it independently checks the algebra the CUDA kernels use.
Reference citations are file:line relative to the reference repository.
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------------------------------
# search space (layers.py:143-152 embed, 425/436-462 attention, 813-821 mlp)
# --------------------------------------------------------------------------------------------------------------------
def embed_widths(D: int) -> List[int]:
    ratios = [i / D for i in range(D // 2, D + 1, min(D // 32, 12))]
    return [int(r * D) for r in ratios]


def head_counts(H: int) -> List[int]:
    return list(range(2, H + 1, 2))


def head_channel_widths(d: int) -> List[int]:
    ratios = [i / d for i in range(d // 4, d + 1, max(d // 8, 1))]
    return [int(d * r) for r in ratios]


def hidden_widths(h: int) -> List[int]:
    ratios = [i / h for i in range(h // 4, h + 1, h // 8)]
    return [int(r * h) for r in ratios]


@dataclass
class ModelCfg:
    embed_dim: int = 384
    num_heads: int = 6
    depth: int = 12
    mlp_ratio: int = 4
    num_classes: int = 1000
    patch: int = 16
    img: int = 224
    eps: float = 1e-6
    # loss weights: search.py:173-179 defaults
    w_attn: float = 0.5
    w_mlp: float = 0.5
    w_embed: float = 0.5
    w_flops: float = 5.0
    target_flops: float = 1.0
    smoothing: float = 0.1

    @property
    def head_dim(self):
        return self.embed_dim // self.num_heads

    @property
    def hidden(self):
        return self.embed_dim * self.mlp_ratio

    @property
    def num_patches(self):
        return (self.img // self.patch) ** 2


@dataclass
class PrunedShape:
    """Shapes and remaining search space of a model that compress() has physically truncated (vision_transformer.py:785-950)
    while every module is still being searched. widths / head_counts are the surviving prefixes of the search-space lists."""
    embed: int                                   # D'
    heads: List[int]                             # H'_l
    head_dims: List[int]                         # d'_l
    hiddens: List[int]                           # h'_l
    spaces: Dict[str, tuple]                     # prefix -> (widths, head_counts)


def w_p_schedule(epoch_frac: float, warmup_epochs: int = 20, hi: float = 0.99, lo: float = 0.1) -> float:
    """layers.py:484-486 update_w."""
    e = min(epoch_frac, warmup_epochs)
    return (lo - hi) / warmup_epochs * e + hi


def keep_ratio_schedule(epoch_frac: float, warmup_epochs: int = 20, max_ratio=0.95, min_ratio=0.75) -> float:
    """vision_transformer.py:521-523 adjust_masking_ratio."""
    e = min(epoch_frac, warmup_epochs)
    return max_ratio - (max_ratio - min_ratio) * e / warmup_epochs


# --------------------------------------------------------------------------------------------------------------------
# bi-mask gate
# --------------------------------------------------------------------------------------------------------------------
def _alive_softmax(alpha: torch.Tensor, switch: torch.Tensor) -> torch.Tensor:
    """softmax over alive cells, zeros elsewhere (layers.py:494-495)."""
    a = torch.where(switch, alpha, torch.full_like(alpha, float("-inf")))
    return torch.softmax(a.reshape(-1), dim=0).reshape(alpha.shape)


def _desc_rank(x: torch.Tensor) -> torch.Tensor:
    """rank[..., c] = position of x[..., c] in a descending stable sort along the last dim (ties: lower index first)."""
    gt = (x.unsqueeze(-1) < x.unsqueeze(-2)).sum(-1)  # how many are strictly greater
    n = x.shape[-1]
    idx = torch.arange(n, device=x.device)
    eq_before = ((x.unsqueeze(-1) == x.unsqueeze(-2)) & (idx.unsqueeze(0) < idx.unsqueeze(1))).sum(-1)
    return gt + eq_before


def gate_1d(alpha, switch, score, widths: List[int], w_p: float):
    """MLP / embed gate (layers.py:847-858, 179-191).  alpha [1,n], score [1,dim].
    returns gate [dim], wr [dim] (= weight_restore), wsum (= weighted_mask.sum())."""
    dim = score.shape[-1]
    if int(switch.sum()) == 1:
        # finished module (layers.py:196-197, 859-860): the frozen gate is the (finalised) score itself, the weighted mask is
        # the all-ones mask of the single surviving cell
        g = score.reshape(-1)
        return g, torch.ones_like(g), torch.tensor(float(dim), dtype=score.dtype, device=score.device)
    a = _alive_softmax(alpha, switch).reshape(-1)
    w = torch.tensor(widths, device=alpha.device)
    r = torch.arange(dim, device=alpha.device)
    # W[r] = sum_j a_j [w_j > r]
    table = (a.unsqueeze(1) * (w.unsqueeze(1) > r.unsqueeze(0)).to(a.dtype)).sum(0)
    rank = _desc_rank(score.reshape(-1))
    wr = table[rank]
    gate = w_p * torch.sigmoid(score.reshape(-1)) + (1 - w_p) * wr
    return gate, wr, table.sum()


def gate_attn(alpha, switch, score, heads: List[int], widths: List[int], w_p: float):
    """joint head x channel gate (layers.py:494-509).  alpha [nh, nc], score [H, d].
    returns gate [H,d], wr [H,d], wsum."""
    H, d = score.shape
    if int(switch.sum()) == 1:                  # finished module (layers.py:518-521): q, k, v *= score
        return score, torch.ones_like(score), torch.tensor(float(H * d), dtype=score.dtype, device=score.device)
    a = _alive_softmax(alpha, switch)
    n_i = torch.tensor(heads, device=alpha.device)
    w_j = torch.tensor(widths, device=alpha.device)
    hr = torch.arange(H, device=alpha.device)
    cr = torch.arange(d, device=alpha.device)
    hm = (n_i.unsqueeze(1) > hr.unsqueeze(0)).to(a.dtype)  # [nh, H]
    cm = (w_j.unsqueeze(1) > cr.unsqueeze(0)).to(a.dtype)  # [nc, d]
    table = hm.t() @ a @ cm                                # [H, d] : W[h_rank, c_rank]
    sig = torch.sigmoid(score)
    rank_c = _desc_rank(score)                             # within head
    rank_h = _desc_rank(sig.sum(-1))                       # heads by sum of sigmoid
    wr = table[rank_h.unsqueeze(1), rank_c]
    gate = w_p * sig + (1 - w_p) * wr
    return gate, wr, table.sum()


def keep_index_sets(score: torch.Tensor, widths: List[int], heads: List[int] = ()):
    """Unit index sets the reference's compress() keeps when it slices a module to a candidate of its search space:
    channels: `torch.argsort(score, descending=True)[:width]` per head (layers.py:614-620, 666-670 attention; 932-933,
    967-968 MLP; 268-269, 308-309 embed); heads: `torch.argsort(sigmoid(score).sum(-1), descending=True)[:n]`.
    score [H, d] (H = 1 for MLP / embed). Returned sorted, like the engine's pruned_index_sets()."""
    score = score.reshape(-1, score.shape[-1])
    order = torch.argsort(score, dim=-1, descending=True, stable=True)
    horder = torch.argsort(torch.sigmoid(score).sum(-1), descending=True, stable=True)
    return {"channels": {int(w): [sorted(order[h, :w].tolist()) for h in range(score.shape[0])] for w in widths},
            "heads": {int(n): sorted(horder[:n].tolist()) for n in heads}}


# --------------------------------------------------------------------------------------------------------------------
# PMIM helpers
# --------------------------------------------------------------------------------------------------------------------
def pmim_mask(noise: torch.Tensor, keep: int) -> torch.Tensor:
    """vision_transformer.py:597-607: 1 = removed.  A patch is kept iff its noise is among the `keep` smallest."""
    rank = _desc_rank(-noise)  # ascending rank
    return (rank >= keep).to(noise.dtype)


def norm_targets(img: torch.Tensor, k: int = 47) -> torch.Tensor:
    """vision_transformer.py:121-141 via box sums over the clipped window (count_include_pad=False)."""
    B, C, H, W = img.shape
    r = k // 2

    def box(x):
        p = F.pad(x, (r + 1, r, r + 1, r))
        c = p.cumsum(-1).cumsum(-2)
        return c[..., k:, k:] - c[..., :-k, k:] - c[..., k:, :-k] + c[..., :-k, :-k]

    x = img.double()
    cnt = box(torch.ones(1, 1, H, W, dtype=torch.float64, device=img.device))
    mean = box(x) / cnt
    sq = box(x * x) / cnt
    var = ((sq - mean * mean) * (cnt / (cnt - 1))).clamp(min=0.)
    return ((x - mean) / (var + 1e-6).sqrt()).to(img.dtype)


def patchify_pixel_shuffle(t: torch.Tensor, P: int = 16) -> torch.Tensor:
    """[B,3,H,W] -> [B, L, 3*P*P] with column c*P*P + i*P + j: the layout a 1x1 conv + PixelShuffle(P) emits per
    patch (vision_transformer.py:491-496, 723)."""
    B, C, H, W = t.shape
    h, w = H // P, W // P
    return t.reshape(B, C, h, P, w, P).permute(0, 2, 4, 1, 3, 5).reshape(B, h * w, C * P * P)


# --------------------------------------------------------------------------------------------------------------------
# Mixup / CutMix of the post-search phase (search.py:651-655, engine.py:98-99 `samples, targets = mixup_fn(samples, targets)`)
# timm.data.Mixup is third-party code that is NOT under /root/reference (requirements.txt:4, an unpinned timm fork; timm is
# not installed here either): this restates the published timm-0.4 `mode='batch'` algorithm (timm/data/mixup.py: Mixup.
# _params_per_batch, _mix_batch, mixup_target, rand_bbox, cutmix_bbox_and_lam with correct_lam=True). PARITY UNPINNED for
# this helper - no reference-side golden exists; the search-step goldens take the mixed batch as their input.
# --------------------------------------------------------------------------------------------------------------------
def mixup_draw(rng, img_hw=(224, 224), mixup_alpha=0.8, cutmix_alpha=1.0, prob=1.0, switch_prob=0.5):
    """One batch-mode draw with a numpy RandomState-like `rng`, in timm's call order: (lam, box or None); box = (yl, yh, xl, xh)
    with lam corrected to the clipped box area."""
    import numpy as np
    lam, use_cutmix = 1.0, False
    if rng.rand() < prob:
        if mixup_alpha > 0. and cutmix_alpha > 0.:
            use_cutmix = rng.rand() < switch_prob
            lam = float(rng.beta(cutmix_alpha, cutmix_alpha) if use_cutmix else rng.beta(mixup_alpha, mixup_alpha))
        elif mixup_alpha > 0.:
            lam = float(rng.beta(mixup_alpha, mixup_alpha))
        elif cutmix_alpha > 0.:
            use_cutmix, lam = True, float(rng.beta(cutmix_alpha, cutmix_alpha))
    if lam == 1.0 or not use_cutmix:
        return lam, None
    H, W = img_hw
    ratio = np.sqrt(1 - lam)
    cut_h, cut_w = int(H * ratio), int(W * ratio)
    cy, cx = rng.randint(0, H), rng.randint(0, W)
    yl, yh = int(np.clip(cy - cut_h // 2, 0, H)), int(np.clip(cy + cut_h // 2, 0, H))
    xl, xh = int(np.clip(cx - cut_w // 2, 0, W)), int(np.clip(cx + cut_w // 2, 0, W))
    lam = 1. - (yh - yl) * (xh - xl) / float(H * W)
    return lam, (yl, yh, xl, xh)


def mixup_target(labels: torch.Tensor, num_classes: int, lam: float, smoothing: float) -> torch.Tensor:
    off = smoothing / num_classes
    on = 1. - smoothing + off
    y1 = torch.full((labels.shape[0], num_classes), off).scatter_(1, labels.view(-1, 1), on)
    y2 = torch.full((labels.shape[0], num_classes), off).scatter_(1, labels.flip(0).view(-1, 1), on)
    return y1 * lam + y2 * (1. - lam)


def mixup_batch(images: torch.Tensor, labels: torch.Tensor, lam: float, box, num_classes: int = 1000, smoothing: float = 0.1):
    """(mixed images, soft targets): x <- lam x + (1 - lam) x.flip(0), or the CutMix box pasted from x.flip(0)."""
    x = images.clone()
    if lam != 1.0:
        if box is not None:
            yl, yh, xl, xh = box
            x[:, :, yl:yh, xl:xh] = images.flip(0)[:, :, yl:yh, xl:xh]
        else:
            x = images * lam + images.flip(0) * (1. - lam)
    return x, mixup_target(labels, num_classes, lam, smoothing)


# --------------------------------------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------------------------------------
@dataclass
class StepInputs:
    images: torch.Tensor            # [B,3,224,224]
    labels: torch.Tensor            # [B] int64
    noise: torch.Tensor             # [B,196] PMIM noise (vision_transformer.py:597)
    drop_scale: torch.Tensor        # [depth, 2, B] DropPath multipliers floor(keep+u)/keep (1.0 where drop prob = 0)
    w_p: float = 0.99
    keep_ratio: float = 0.95
    soft_target: Optional[torch.Tensor] = None   # [B, C] Mixup targets of the post-search phase (search.py:651-655)


@dataclass
class StepOutputs:
    logits: torch.Tensor
    loss_base: torch.Tensor
    loss_arch: torch.Tensor
    loss_decoder: torch.Tensor
    loss_total: torch.Tensor
    loss_terms: Dict[str, torch.Tensor] = field(default_factory=dict)
    mask: Optional[torch.Tensor] = None
    gates: Dict[str, torch.Tensor] = field(default_factory=dict)


def default_switches(cfg: ModelCfg):
    sw = {"patch_embed": torch.ones(1, len(embed_widths(cfg.embed_dim)), dtype=torch.bool)}
    for l in range(cfg.depth):
        sw[f"blocks.{l}.attn"] = torch.ones(len(head_counts(cfg.num_heads)), len(head_channel_widths(cfg.head_dim)),
                                            dtype=torch.bool)
        sw[f"blocks.{l}.mlp"] = torch.ones(1, len(hidden_widths(cfg.hidden)), dtype=torch.bool)
    return sw


def _ln(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def _sparsity_term(alpha, switch, score, coef):
    """base_model.py:58-78 for one searchable module (entropy + variance-tan + score norm)."""
    p = torch.softmax(alpha[switch], dim=-1)
    n = int(switch.sum())
    loss = -(p * p.log()).sum()
    sigma = ((p - p.mean()) ** 2).sum() / (1. - 1. / n)
    loss = loss + torch.tan(math.pi / 2 - math.pi * sigma) / n
    loss = loss + torch.sigmoid(score).sum() * coef
    return loss


def forward_step(P: Dict[str, torch.Tensor], inp: StepInputs, cfg: ModelCfg, switches=None, shape: Optional[PrunedShape] = None,
                 finish_search: bool = False) -> StepOutputs:
    """MIMVisionTransformer.forward (vision_transformer.py:614-669, 717-745) in search/training mode with embed,
    attention and MLP search active, followed by OFBSearchLOSS (losses.py:80-106) and the decoder-loss weighting of
    engine.search_one_epoch (engine.py:131-144).
    shape: the model after truncating prune events (tensors physically sliced, every module still searched). The attention
    scale and the ORIGINAL-FLOPs side of the FLOPs loss keep the unpruned dims (layers.py:418, 747-753); the
    searched side uses the pruned LayerNorm width and head counts (vt:206-213 active_dim, layers.py:749 active_H).
    finish_search: every module is finalised - the criterion returns the base loss alone (losses.py:105-106), there is no
    architecture optimizer any more (engine.py:206-208). With inp.keep_ratio == 1 (reset_mask_ratio(1.0), search.py:645) the
    PMIM branch is off (vt:595-612, 719); inp.soft_target switches the base criterion to timm SoftTargetCrossEntropy on Mixup
    targets (search.py:651-655)."""
    sw = switches or default_switches(cfg)
    D, H, d, hid, L = cfg.embed_dim, cfg.num_heads, cfg.head_dim, cfg.hidden, cfg.num_patches
    Dc = shape.embed if shape is not None else D                      # current (pruned) embedding width
    sp = (lambda k, default: shape.spaces[k] if shape is not None else default)
    B = inp.images.shape[0]
    w_p = inp.w_p
    gates = {}

    # ---- patch embed + embed gate (layers.py:173-191) ----
    g_e, wr_e, wsum_e = gate_1d(P["patch_embed.alpha"], sw["patch_embed"], P["patch_embed.score"],
                                sp("patch_embed", (embed_widths(D), []))[0], w_p)
    gates["patch_embed"] = g_e
    patches = inp.images.reshape(B, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(B, L, 768)
    x = patches @ P["patch_embed.proj.weight"].reshape(Dc, 768).t() + P["patch_embed.proj.bias"]
    x = x * g_e
    # pos embed, PMIM masking, mask token, cls (vision_transformer.py:628-651); masked patches lose their pos-embed
    x = x + P["pos_embed"][0, 1:] * g_e
    keep = int(L * inp.keep_ratio)
    mask = pmim_mask(inp.noise, keep) if keep != L else None
    if mask is not None:
        x = x * (1 - mask).unsqueeze(-1) + mask.unsqueeze(-1) * (P["mask_token"].reshape(1, 1, Dc) * g_e)
    cls = ((P["cls_token"] + P["pos_embed"][:, :1]) * g_e).expand(B, -1, -1)
    x = torch.cat([cls, x], dim=1)
    N = L + 1

    # ---- blocks: search-mode normalised residual stream (vision_transformer.py:193-201) ----
    scale = d ** -0.5
    # once the embedding search is finished the weighted embed mask has no fractional entry and MAEBlock takes its standard
    # pre-norm branch (vision_transformer.py:193, 203-204); until then the residual is taken from the NORMALISED x (194-201)
    prenorm = int(sw["patch_embed"].sum()) == 1
    attn_terms, mlp_terms, attn_wsum, mlp_wsum = [], [], [], []
    for l in range(cfg.depth):
        pre = f"blocks.{l}."
        x_in = x
        x = _ln(x, P[pre + "norm1.weight"], P[pre + "norm1.bias"], cfg.eps)
        Hl, dl = (shape.heads[l], shape.head_dims[l]) if shape is not None else (H, d)
        wj_a, ni_a = sp(pre + "attn", (head_channel_widths(d), head_counts(H)))
        g_a, _, ws_a = gate_attn(P[pre + "attn.alpha"], sw[pre + "attn"], P[pre + "attn.score"], ni_a, wj_a, w_p)
        gates[pre + "attn"] = g_a
        qkv = x @ P[pre + "attn.qkv.weight"].t() + P[pre + "attn.qkv.bias"]
        qkv = qkv.reshape(B, N, 3, Hl, dl) * g_a                    # q,k,v all gated (layers.py:507-509)
        q, k, v = qkv.permute(2, 0, 3, 1, 4)
        att = torch.softmax((q @ k.transpose(-2, -1)) * scale, dim=-1)
        o = (att @ v).transpose(1, 2).reshape(B, N, Hl * dl)
        o = o @ P[pre + "attn.proj.weight"].t() + P[pre + "attn.proj.bias"]
        x = (x_in if prenorm else x) + inp.drop_scale[l, 0].reshape(B, 1, 1) * o
        x_mid = x
        x = _ln(x, P[pre + "norm2.weight"], P[pre + "norm2.bias"], cfg.eps)
        g_m, _, ws_m = gate_1d(P[pre + "mlp.alpha"], sw[pre + "mlp"], P[pre + "mlp.score"],
                               sp(pre + "mlp", (hidden_widths(hid), []))[0], w_p)
        gates[pre + "mlp"] = g_m
        hdn = F.gelu((x @ P[pre + "mlp.fc1.weight"].t() + P[pre + "mlp.fc1.bias"]) * g_m)
        y = hdn @ P[pre + "mlp.fc2.weight"].t() + P[pre + "mlp.fc2.bias"]
        x = (x_mid if prenorm else x) + inp.drop_scale[l, 1].reshape(B, 1, 1) * y
        attn_wsum.append(ws_a)
        mlp_wsum.append(ws_m)
        if int(sw[pre + "attn"].sum()) > 1:
            attn_terms.append(_sparsity_term(P[pre + "attn.alpha"], sw[pre + "attn"], P[pre + "attn.score"], 4e-4))
        if int(sw[pre + "mlp"].sum()) > 1:
            mlp_terms.append(_sparsity_term(P[pre + "mlp.alpha"], sw[pre + "mlp"], P[pre + "mlp.score"], 1e-4))
    latent = _ln(x, P["norm.weight"], P["norm.bias"], cfg.eps)

    # ---- PMIM decoder branch (vision_transformer.py:720-729) ----
    if mask is not None:
        rec = latent[:, 1:] @ P["decoder.0.weight"].reshape(768, Dc).t() + P["decoder.0.bias"]   # [B,L,768]
        tgt = patchify_pixel_shuffle(norm_targets(inp.images, 47))
        l1 = (tgt - rec).abs() * mask.unsqueeze(-1)
        loss_dec = l1.sum() / (mask.sum() * 256 + 1e-5) / 3
    else:
        loss_dec = torch.zeros((), dtype=x.dtype)

    logits = latent[:, 0] @ P["head.weight"].t() + P["head.bias"]

    # ---- losses ----
    logp = F.log_softmax(logits, dim=-1)
    nll = -logp.gather(1, inp.labels.unsqueeze(1)).squeeze(1)
    if inp.soft_target is not None:
        loss_base = (-inp.soft_target * logp).sum(-1).mean()        # timm SoftTargetCrossEntropy
    else:
        loss_base = ((1 - cfg.smoothing) * nll + cfg.smoothing * (-logp.mean(-1))).mean()

    zero = torch.zeros((), dtype=x.dtype)
    l_attn = sum(attn_terms) if attn_terms else zero
    l_mlp = sum(mlp_terms) if mlp_terms else zero
    l_embed = _sparsity_term(P["patch_embed.alpha"], sw["patch_embed"], P["patch_embed.score"], 1e-4) \
        if int(sw["patch_embed"].sum()) > 1 else zero

    # FLOPs model (vision_transformer.py:759-783; layers.py:747-766, 1032-1044), n = N = 196 patches
    n = float(L)
    ae = wsum_e
    f_ori = L * D * 768.
    f_s = L * ae * 768.
    for l in range(cfg.depth):
        sd, sm = attn_wsum[l], mlp_wsum[l]
        Ha = shape.heads[l] if shape is not None else H               # active_H (layers.py:749)
        f_ori += 2 * D * n
        f_s = f_s + 2 * Dc * n                                        # active_dim = norm1.normalized_shape[0] (vt:210)
        f_ori += n * (D * 3 * D) + 3 * n * D + H * n * d * n + H * n * n + 5 * H * n * n + H * n * n * d + n * D * D + n * D
        f_s = f_s + n * (ae * 3 * sd) + 3 * n * sd + n * n * sd + Ha * n * n + 5 * Ha * n * n + n * n * sd \
            + n * (sd * ae) + n * ae
        f_ori += (2 * D * hid + D + hid) * n
        f_s = f_s + (ae * sm * 2 + ae + sm) * n
    f_ori += D * cfg.num_classes
    f_s = f_s + ae * cfg.num_classes
    l_flops = ((f_s / 1e9 - cfg.target_flops) / (f_ori / 1e9)) ** 2

    loss_arch = cfg.w_attn * l_attn + cfg.w_mlp * l_mlp + cfg.w_embed * l_embed + cfg.w_flops * l_flops
    loss_total = loss_base if finish_search else loss_base + loss_arch
    if mask is not None:
        w_dec = (loss_base / loss_dec).detach()          # engine.py:140-142
        loss_total = loss_total + w_dec * loss_dec
    return StepOutputs(logits=logits, loss_base=loss_base, loss_arch=loss_arch, loss_decoder=loss_dec,
                       loss_total=loss_total,
                       loss_terms={"attn": l_attn, "mlp": l_mlp, "embed": l_embed, "flops": l_flops,
                                   "flops_searched": f_s / 1e9, "flops_ori": torch.tensor(f_ori / 1e9)},
                       mask=mask, gates=gates)


# --------------------------------------------------------------------------------------------------------------------
# optimizer (optim.py:56-120) and parameter grouping (search.py:486-559)
# --------------------------------------------------------------------------------------------------------------------
NO_DECAY_KEYS = ("pos_embed", "cls_token", "dist_token", "scale_weight", "mask_token", "score")


def param_group(name: str, p: torch.Tensor) -> str:
    """'param_nd' | 'param_d' | 'dec_nd' | 'dec_d' | 'arch'  (search.py:489-508)."""
    if p.dim() == 1 or name.endswith(".bias") or any(k in name for k in NO_DECAY_KEYS):
        return "dec_nd" if "decoder" in name else "param_nd"
    if "alpha" in name:
        return "arch"
    return "dec_d" if "decoder" in name else "param_d"


def group_hparams(group: str, lr: float, wd: float = 1e-3):
    betas = (0.5, 0.999) if group == "arch" else (0.9, 0.999)
    decay = 0.0 if group.endswith("_nd") else wd
    return dict(lr=lr, betas=betas, eps=1e-8, weight_decay=decay)


def adamw_step(p, g, m, v, step, lr, betas, eps, weight_decay):
    """optim.py:74-118: decay first, then Adam with bias-corrected denominator. In-place on p, m, v."""
    b1, b2 = betas
    p.mul_(1 - lr * weight_decay)
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    denom = (v.sqrt() / math.sqrt(1 - b2 ** step)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / (1 - b1 ** step)))


def train_step(P: Dict[str, torch.Tensor], state: Dict[str, Dict[str, torch.Tensor]], inp: StepInputs, cfg: ModelCfg,
               lr: float, step: int, switches=None, frozen=("alpha_patch",), shape: Optional[PrunedShape] = None,
               finish_search: bool = False):
    """One full search step: forward, losses, backward, three AdamW updates (engine.py:131-184).
    P entries are leaf tensors; returns (outputs, grads). `frozen` names get no gradient and no update (alpha_patch; in the
    post-search phase mask_token and the decoder, vision_transformer.py:534-539; with finish_search the alphas, whose
    optimizer is gone)."""
    leaves = {k: (v.detach().clone().requires_grad_(True) if k not in frozen else v.detach()) for k, v in P.items()}
    out = forward_step(leaves, inp, cfg, switches, shape, finish_search)
    out.loss_total.backward()
    grads = {k: (v.grad if v.grad is not None else None) for k, v in leaves.items() if k not in frozen}
    with torch.no_grad():
        for k, g in grads.items():
            if g is None:
                continue
            st = state.setdefault(k, {"m": torch.zeros_like(P[k]), "v": torch.zeros_like(P[k])})
            hp = group_hparams(param_group(k, P[k]), lr)
            adamw_step(P[k], g, st["m"], st["v"], step, **hp)
    return out, grads
