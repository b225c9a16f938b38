"""Prune event of the search: the decision and the index sets of `compress()`, and the gather of every
parameter tensor into its physically pruned shape.

Reference: MIMVisionTransformer.compress (vision_transformer.py:785-950) -> MAEPatchEmbed.compress (layers.py:218-338),
MAESparseAttention.compress / compress_patchembed (layers.py:559-728), MAESparseMlp.compress / compress_patchembed
(layers.py:883-1025). Per searchable module, with alpha averaged over the ranks (layers.py:9-14):
    p = softmax(alpha[alive]);  thr = thresh / n_alive;   nothing happens unless min p <= thr
    switch' = softmax(alpha | alive) > thr ;  alpha' = alpha where switch' else 0        (the optimizer state of alpha restarts)
    one cell left            -> FINALISE: slice to that cell's (heads, width); score' = w_p sigmoid(score)[kept] + (1 - w_p)
    last row / column dead   -> TRUNCATE: slice alpha / switch to the largest alive (row, column) and the tensors to its sizes
    otherwise                -> only switch' / alpha'
    kept units: heads  = argsort(sigmoid(score).sum(-1), descending)[:heads]         (rank order, NOT sorted)
                channels = argsort(score, descending)[:, :width] of the kept heads   (rank order)
The ranks come from the bi-mask forward kernel (`BimaskTable.rank`, ties -> lower index), so the index sets are the ones the
gates of the last step were built from. What this module does NOT do yet: rebuild the search engine on the truncated shapes
(post-prune search steps) and the Adam-state surgery of optim.AdamW.update (optim.py:122-182); `SearchStepEngine.apply_prune`
therefore applies switch-only events in place and refuses truncating ones. The gathered tensors are what
`FinetuneStepEngine` consumes once every module is finalised.
"""
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch


@dataclass
class ModulePlan:
    prefix: str
    kind: int                      # 0 embed, 1 mlp, 2 attention
    executed: bool                 # execute_prune of the reference
    finished: bool                 # finish_search after the event
    switch: torch.Tensor           # bool [n_i', n_j'] (already truncated)
    alpha: torch.Tensor            # fp32 [n_i', n_j']
    truncated: bool = False        # tensors are physically sliced (TRUNCATE or FINALISE)
    finalised: bool = False
    head_num: int = 0              # attention: heads kept
    width: int = 0                 # channels per head / hidden units / embedding dims kept
    head_index: Optional[torch.Tensor] = None      # [head_num] original head ids in rank order
    channel_index: Optional[torch.Tensor] = None   # [head_num or 1, width] original channel ids in rank order


def _alive_softmax(alpha, switch):
    a = torch.where(switch, alpha, torch.full_like(alpha, float("-inf")))
    return torch.softmax(a.reshape(-1), dim=0).reshape(alpha.shape)


def plan_module(prefix: str, kind: int, alpha: torch.Tensor, switch: torch.Tensor, widths: List[int], head_counts: List[int],
                head_rank: torch.Tensor, chan_rank: torch.Tensor, thresh: float) -> ModulePlan:
    """alpha / switch [n_i, n_j] (n_i = 1 for embed / mlp); head_rank [H], chan_rank [H, dim]: descending-order ranks of the
    last forward. All tensors on the CPU (a prune event happens three times per epoch)."""
    alpha, switch = alpha.detach().float().cpu(), switch.detach().bool().cpu()
    n_alive = int(switch.sum())
    if n_alive == 1:                                                   # layers.py:560-563
        return ModulePlan(prefix, kind, False, True, switch, alpha)
    p = torch.softmax(alpha[switch].reshape(-1), dim=0)
    thr = thresh / n_alive
    if float(p.min()) > thr:                                           # layers.py:574 / 696
        return ModulePlan(prefix, kind, False, False, switch, alpha)
    new_switch = _alive_softmax(alpha, switch) > thr                   # layers.py:578-581
    new_alpha = torch.where(new_switch, alpha, torch.zeros_like(alpha))
    plan = ModulePlan(prefix, kind, True, False, new_switch, new_alpha)
    idx = torch.nonzero(new_switch)
    H = head_rank.numel()
    if idx.shape[0] == 1:                                              # FINALISE (layers.py:597-645, 925-952, 261-292)
        i, j = int(idx[0, 0]), int(idx[0, 1])
        plan.finished = plan.finalised = plan.truncated = True
    elif kind == 2 and (int(new_switch[:, -1].sum()) == 0 or int(new_switch[-1, :].sum()) == 0):     # layers.py:647
        i, j = int(idx[:, 0].max()), int(idx[:, 1].max())
        plan.truncated = True
    elif kind != 2 and int(new_switch[0, -1]) == 0:                    # layers.py:954, 294
        i, j = 0, int(idx[-1, 1])
        plan.truncated = True
    else:
        return plan
    plan.head_num = head_counts[i] if kind == 2 else 1
    plan.width = widths[j]
    if not plan.finalised:
        plan.switch = new_switch[:i + 1, :j + 1].clone()
        plan.alpha = new_alpha[:i + 1, :j + 1].clone()
    head_order = torch.argsort(head_rank.long().cpu()) if H > 1 else torch.arange(1)
    plan.head_index = head_order[:plan.head_num].clone()
    chan_order = torch.argsort(chan_rank.long().cpu(), dim=-1)        # [H, dim]: original channel id at each rank
    plan.channel_index = chan_order[plan.head_index][:, :plan.width].clone()
    return plan


def qkv_keep_index(plan: ModulePlan, H: int, d: int) -> torch.Tensor:
    """Rows of qkv.weight / qkv.bias kept by an attention plan (layers.py:615-620)."""
    base = torch.arange(3).view(3, 1, 1) * (H * d) + plan.head_index.view(1, -1, 1) * d
    return (base + plan.channel_index.unsqueeze(0)).reshape(-1)


def proj_keep_index(plan: ModulePlan, d: int) -> torch.Tensor:
    """Input columns of attn.proj.weight kept by an attention plan (layers.py:636-637)."""
    return (plan.head_index.view(-1, 1) * d + plan.channel_index).reshape(-1)


def gather_pruned(plans: Dict[str, ModulePlan], named: Dict[str, torch.Tensor], dims: Dict[str, dict], w_p: float,
                  state: bool = False):
    """Every tensor of `named` (reference state_dict names, reference shapes, any device) in the shape the reference's
    compress() leaves it in. dims[prefix] = {"heads": H, "dim": d} of the modules BEFORE the event.
    state=True: `named` holds an Adam moment (exp_avg or exp_avg_sq) instead of the parameters; optim.AdamW.update
    (optim.py:122-182) slices it with the same indices, except that the alpha of a module that executed a prune and a finalised
    score restart from zero (initialize=True)."""
    out = {k: v for k, v in named.items()}
    dev = next(iter(named.values())).device
    if state:
        out_state = gather_pruned(plans, named, dims, w_p, state=False)
        for prefix, pl in plans.items():
            n_i, n_j = pl.switch.shape
            a = named[prefix + ".alpha"]
            out_state[prefix + ".alpha"] = torch.zeros(n_i, n_j, dtype=a.dtype, device=dev) if pl.executed \
                else a.reshape(-1, a.shape[-1])[:n_i, :n_j].clone()
            if pl.finalised:
                out_state[prefix + ".score"] = torch.zeros_like(out_state[prefix + ".score"])
        return out_state

    def sel(name, index, axis):
        out[name] = out[name].index_select(axis, index.to(dev))

    pe = plans.get("patch_embed")
    keep_e = None
    if pe is not None and pe.truncated:                                # layers.py:261-338 + vt:837-915
        keep_e = pe.channel_index.reshape(-1)
        score = named["patch_embed.score"]
        if pe.finalised:
            out["patch_embed.score"] = w_p * torch.sigmoid(score).index_select(-1, keep_e.to(dev)) + (1 - w_p)
        else:
            sel("patch_embed.score", keep_e, 1)
        out["patch_embed.proj.weight"] = named["patch_embed.proj.weight"].index_select(0, keep_e.to(dev))
        sel("patch_embed.proj.bias", keep_e, 0)
        for k in ("mask_token", "cls_token", "pos_embed"):
            sel(k, keep_e, 2)
        sel("norm.weight", keep_e, 0); sel("norm.bias", keep_e, 0)
        sel("head.weight", keep_e, 1)
        sel("decoder.0.weight", keep_e, 1)
    for prefix, pl in plans.items():
        if prefix == "patch_embed":
            out[prefix + ".alpha"] = pl.alpha.to(dev)
            continue
        out[prefix + ".alpha"] = pl.alpha.to(dev)
        blk = prefix.rsplit(".", 1)[0]
        if pl.kind == 2:
            if pl.truncated:
                H, d = dims[prefix]["heads"], dims[prefix]["dim"]
                score = named[prefix + ".score"]
                kept = score.index_select(0, pl.head_index.to(dev)).gather(1, pl.channel_index.to(dev))
                out[prefix + ".score"] = (w_p * torch.sigmoid(kept) + (1 - w_p)) if pl.finalised else kept
                kq = qkv_keep_index(pl, H, d)
                sel(prefix + ".qkv.weight", kq, 0); sel(prefix + ".qkv.bias", kq, 0)
                sel(prefix + ".proj.weight", proj_keep_index(pl, d), 1)
            if keep_e is not None:                                     # compress_patchembed layers.py:698-711
                sel(prefix + ".qkv.weight", keep_e, 1)
                sel(prefix + ".proj.weight", keep_e, 0); sel(prefix + ".proj.bias", keep_e, 0)
                sel(blk + ".norm1.weight", keep_e, 0); sel(blk + ".norm1.bias", keep_e, 0)
                sel(blk + ".norm2.weight", keep_e, 0); sel(blk + ".norm2.bias", keep_e, 0)
        else:
            if pl.truncated:
                kc = pl.channel_index.reshape(-1)
                score = named[prefix + ".score"]
                if pl.finalised:
                    out[prefix + ".score"] = w_p * torch.sigmoid(score).index_select(-1, kc.to(dev)) + (1 - w_p)
                else:
                    sel(prefix + ".score", kc, 1)
                sel(prefix + ".fc1.weight", kc, 0); sel(prefix + ".fc1.bias", kc, 0)
                sel(prefix + ".fc2.weight", kc, 1)
            if keep_e is not None:                                     # layers.py:994-1008
                sel(prefix + ".fc1.weight", keep_e, 1)
                sel(prefix + ".fc2.weight", keep_e, 0); sel(prefix + ".fc2.bias", keep_e, 0)
    if pe is not None:
        out["patch_embed.alpha"] = pe.alpha.to(dev)
    return out


def fuse_params(named: Dict[str, torch.Tensor], prefixes: List[str]) -> Dict[str, torch.Tensor]:
    """MIMVisionTransformer.fuse (vision_transformer.py:747-757) + the module fuse() methods (layers.py:202-206, 539-543,
    867-871) on a fully finalised model: the frozen gates (the finalised `score` tensors) are folded into the weights that
    produce the gated activations, which leaves a plain pre-norm ViT on pruned shapes - the model `finetune.intersect` copies
    and `FinetuneStepEngine` trains / evaluates. Returns the tensors under the plain VisionTransformer's names (scores, alphas,
    mask token and decoder dropped)."""
    out = {k: v for k, v in named.items()}
    se = named["patch_embed.score"].reshape(-1)
    for k in ("cls_token", "pos_embed"):
        out[k] = named[k] * se
    out["patch_embed.proj.weight"] = named["patch_embed.proj.weight"] * se.view(-1, 1, 1, 1)
    out["patch_embed.proj.bias"] = named["patch_embed.proj.bias"] * se
    for p in prefixes:
        if p == "patch_embed":
            continue
        s = named[p + ".score"].reshape(-1)
        if p.endswith(".attn"):
            s3 = s.repeat(3)
            out[p + ".qkv.weight"] = named[p + ".qkv.weight"] * s3.unsqueeze(-1)
            out[p + ".qkv.bias"] = named[p + ".qkv.bias"] * s3
        else:
            out[p + ".fc1.weight"] = named[p + ".fc1.weight"] * s.unsqueeze(-1)
            out[p + ".fc1.bias"] = named[p + ".fc1.bias"] * s
    return {k: v for k, v in out.items()
            if not (k.endswith(".score") or k.endswith(".alpha") or k == "mask_token" or k.startswith("decoder."))}


def subnet_dims(plans: Dict[str, ModulePlan], depth: int):
    """(embed_dim, heads[], head_dims[], hiddens[]) of a fully finalised model, the constructor arguments of
    FinetuneStepEngine."""
    assert all(pl.finalised for pl in plans.values()), "every searchable module must be finalised"
    return (plans["patch_embed"].width, [plans[f"blocks.{l}.attn"].head_num for l in range(depth)],
            [plans[f"blocks.{l}.attn"].width for l in range(depth)], [plans[f"blocks.{l}.mlp"].width for l in range(depth)])
