"""TEST INFRASTRUCTURE ONLY — deterministic (seeded, CPU-generated) parameters and inputs shared by
oracle/make_golden.py, tests/ and bench.py's cpu_baseline leg.  Parameter names follow the reference state_dict
(MIMVisionTransformer, vision_transformer.py:385-519).  Init statistics follow the reference (trunc-normal .02 weights,
alpha ~ U(0,1), score ~ trunc-normal .2: layers.py:147-155, 455-467, 817-824) but the head and biases get small random
values instead of zeros so that every gradient path is exercised.
"""
from typing import Dict

import torch

from ofb_oracle import (ModelCfg, StepInputs, embed_widths, head_channel_widths, head_counts, hidden_widths)

SUMMARY_SAMPLES = 48


def make_params(cfg: ModelCfg, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    D, H, d, hid, L, C = cfg.embed_dim, cfg.num_heads, cfg.head_dim, cfg.hidden, cfg.num_patches, cfg.num_classes

    def tn(*shape, std=.02):
        return (torch.randn(*shape, generator=g) * std).clamp_(-2 * std, 2 * std)

    def uni(*shape):
        return torch.rand(*shape, generator=g)

    P = {
        "cls_token": tn(1, 1, D), "pos_embed": tn(1, L + 1, D), "mask_token": tn(1, 1, D),
        "alpha_patch": torch.ones(1, 1),
        "patch_embed.alpha": uni(1, len(embed_widths(D))), "patch_embed.score": tn(1, D, std=.2),
        "patch_embed.proj.weight": tn(D, 3, cfg.patch, cfg.patch, std=.05), "patch_embed.proj.bias": tn(D),
    }
    for l in range(cfg.depth):
        p = f"blocks.{l}."
        P[p + "norm1.weight"] = 1 + tn(D, std=.1)
        P[p + "norm1.bias"] = tn(D, std=.1)
        P[p + "attn.alpha"] = uni(len(head_counts(H)), len(head_channel_widths(d)))
        P[p + "attn.score"] = tn(H, d, std=.2)
        P[p + "attn.qkv.weight"] = tn(3 * D, D, std=.04)
        P[p + "attn.qkv.bias"] = tn(3 * D)
        P[p + "attn.proj.weight"] = tn(D, D, std=.04)
        P[p + "attn.proj.bias"] = tn(D)
        P[p + "norm2.weight"] = 1 + tn(D, std=.1)
        P[p + "norm2.bias"] = tn(D, std=.1)
        P[p + "mlp.alpha"] = uni(1, len(hidden_widths(hid)))
        P[p + "mlp.score"] = tn(1, hid, std=.2)
        P[p + "mlp.fc1.weight"] = tn(hid, D, std=.04)
        P[p + "mlp.fc1.bias"] = tn(hid)
        P[p + "mlp.fc2.weight"] = tn(D, hid, std=.04)
        P[p + "mlp.fc2.bias"] = tn(D)
    P["norm.weight"] = 1 + tn(D, std=.1)
    P["norm.bias"] = tn(D, std=.1)
    P["head.weight"] = tn(C, D, std=.05)
    P["head.bias"] = tn(C)
    P["decoder.0.weight"] = tn(768, D, 1, 1, std=.05)
    P["decoder.0.bias"] = tn(768)
    return P


def make_inputs(cfg: ModelCfg, batch: int, seed: int = 1, epoch_frac: float = 0.0, drop_path_rate: float = 0.1,
                w_p=None, keep_ratio=None) -> StepInputs:
    """Synthetic ImageNet-shaped batch + the random draws of one step (PMIM noise, DropPath)."""
    from ofb_oracle import keep_ratio_schedule, w_p_schedule
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(batch, 3, cfg.img, cfg.img, generator=g)
    # give the images some low-frequency structure so the 47x47 normalisation is not trivially ~N(0,1)
    images = images + 2.0 * torch.nn.functional.interpolate(
        torch.randn(batch, 3, 7, 7, generator=g), size=(cfg.img, cfg.img), mode="bilinear", align_corners=False)
    labels = torch.randint(0, cfg.num_classes, (batch,), generator=g)
    noise = torch.rand(batch, cfg.num_patches, generator=g)
    dpr = torch.linspace(0, drop_path_rate, cfg.depth)
    scale = torch.ones(cfg.depth, 2, batch)
    draws = []   # the U[0,1) draws timm's DropPath would make, in call order (attn then mlp, block by block)
    for l in range(cfg.depth):
        p = float(dpr[l])
        if p > 0:
            keep = 1 - p
            u = torch.rand(2, batch, generator=g)
            scale[l] = torch.floor(keep + u) / keep
            draws += [u[0], u[1]]
    inp = StepInputs(images=images, labels=labels, noise=noise, drop_scale=scale,
                      w_p=w_p if w_p is not None else w_p_schedule(epoch_frac),
                      keep_ratio=keep_ratio if keep_ratio is not None else keep_ratio_schedule(epoch_frac))
    inp.drop_draws = draws
    return inp


def summarize(t: torch.Tensor) -> torch.Tensor:
    """Compact, order-sensitive fingerprint of a tensor: [sum, abs-sum, l2, SUMMARY_SAMPLES strided samples]."""
    f = t.detach().double().reshape(-1)
    n = f.numel()
    idx = (torch.arange(SUMMARY_SAMPLES, dtype=torch.float64) * (n - 1) / max(SUMMARY_SAMPLES - 1, 1)).long()
    return torch.cat([torch.stack([f.sum(), f.abs().sum(), f.pow(2).sum().sqrt()]), f[idx]])


def search_modules(cfg: ModelCfg):
    """(prefix, kind, heads, dim, widths, head_counts) of every searchable module in model order (kind 0 embed, 1 mlp, 2 attn)."""
    mods = [("patch_embed", 0, 1, cfg.embed_dim, embed_widths(cfg.embed_dim), [])]
    for l in range(cfg.depth):
        mods.append((f"blocks.{l}.attn", 2, cfg.num_heads, cfg.head_dim, head_channel_widths(cfg.head_dim),
                     head_counts(cfg.num_heads)))
        mods.append((f"blocks.{l}.mlp", 1, 1, cfg.hidden, hidden_widths(cfg.hidden), []))
    return mods


def pruned_shape_from_plans(cfg: ModelCfg, plans):
    """PrunedShape (oracle) of the model a set of prune plans (ofb_b200.prune.ModulePlan, duck-typed) leaves behind."""
    from ofb_oracle import PrunedShape
    spaces, heads, dims, hids = {}, [], [], []
    embed = cfg.embed_dim
    for prefix, kind, H, dim, widths, counts in search_modules(cfg):
        pl = plans[prefix]
        n_i, n_j = pl.switch.shape
        spaces[prefix] = (list(widths[:n_j]), list(counts[:n_i]))
        if kind == 0:
            embed = pl.width if pl.truncated else dim
        elif kind == 2:
            heads.append(pl.head_num if pl.truncated else H)
            dims.append(pl.width if pl.truncated else dim)
        else:
            hids.append(pl.width if pl.truncated else dim)
    return PrunedShape(embed=embed, heads=heads, head_dims=dims, hiddens=hids, spaces=spaces)
