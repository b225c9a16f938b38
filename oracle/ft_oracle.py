"""TEST INFRASTRUCTURE ONLY — CPU oracle of the *finetune* training step of a physically pruned (searched) subnet
(BASELINE.json configs[4]). Nothing in the product (once-for-both_b200/) may import it.

Reference path restated here (file:line relative to the reference repository):
  * model  : plain VisionTransformer.forward (models/vision_transformer.py:332-358) on per-layer pruned shapes, pre-norm
             Block (vision_transformer.py:157-160), Attention (models/layers.py:382-394: head dim inferred from the sliced qkv,
             scale fixed at construction = (D/H)^-0.5 of the UNPRUNED model, layers.py:375), Mlp (layers.py:784-790),
             PatchEmbed conv (layers.py:121-128); the shapes are what finetune.intersect (finetune.py:182-249) installs.
  * loss   : timm LabelSmoothingCrossEntropy(0.1) or SoftTargetCrossEntropy (finetune.py:388-394) through DistillationLoss
             with distillation 'none' (losses.py:25-40).
  * update : torch.optim.AdamW over lr_decay.param_groups_lrd (lr_decay.py:15-75; finetune.py:378-383): layer-wise lr
             scale layer_decay^(depth+1-layer), no weight decay for 1-D tensors / cls_token / pos_embed.
  * step   : engine.train_one_epoch loop body (engine.py:31-62).

Parity pin: oracle/make_golden_ft.py runs the UNMODIFIED reference classes on identical seeded parameters / inputs and
stores their outputs in tests/golden/ft_*.npz; tests/test_ft_oracle_golden.py re-checks this restatement against them.
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


@dataclass
class SubnetCfg:
    embed_dim: int = 384                 # pruned embedding width D'
    heads: List[int] = field(default_factory=lambda: [6] * 12)        # H'_l
    head_dims: List[int] = field(default_factory=lambda: [64] * 12)   # d'_l
    hiddens: List[int] = field(default_factory=lambda: [1536] * 12)   # h'_l
    num_classes: int = 1000
    patch: int = 16
    img: int = 224
    eps: float = 1e-6
    scale: float = 0.125                 # (D/H)^-0.5 of the unpruned DeiT (64^-0.5), never re-derived
    smoothing: float = 0.1

    @property
    def depth(self):
        return len(self.heads)

    @property
    def num_patches(self):
        return (self.img // self.patch) ** 2


def ft_param_shapes(cfg: SubnetCfg) -> Dict[str, tuple]:
    """state_dict names / shapes of the plain VisionTransformer after finetune.intersect, in named_parameters() order."""
    D, L, C = cfg.embed_dim, cfg.num_patches, cfg.num_classes
    s = {"cls_token": (1, 1, D), "pos_embed": (1, L + 1, D),
         "patch_embed.proj.weight": (D, 3, cfg.patch, cfg.patch), "patch_embed.proj.bias": (D,)}
    for l in range(cfg.depth):
        p, A, hid = f"blocks.{l}.", cfg.heads[l] * cfg.head_dims[l], cfg.hiddens[l]
        s[p + "norm1.weight"] = (D,); s[p + "norm1.bias"] = (D,)
        s[p + "attn.qkv.weight"] = (3 * A, D); s[p + "attn.qkv.bias"] = (3 * A,)
        s[p + "attn.proj.weight"] = (D, A); s[p + "attn.proj.bias"] = (D,)
        s[p + "norm2.weight"] = (D,); s[p + "norm2.bias"] = (D,)
        s[p + "mlp.fc1.weight"] = (hid, D); s[p + "mlp.fc1.bias"] = (hid,)
        s[p + "mlp.fc2.weight"] = (D, hid); s[p + "mlp.fc2.bias"] = (D,)
    s["norm.weight"] = (D,); s["norm.bias"] = (D,)
    s["head.weight"] = (C, D); s["head.bias"] = (C,)
    return s


def make_ft_params(cfg: SubnetCfg, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded parameters with the reference's init statistics but non-zero biases (every gradient path is exercised)."""
    g = torch.Generator().manual_seed(seed)
    P = {}
    for k, shp in ft_param_shapes(cfg).items():
        if k.endswith("norm1.weight") or k.endswith("norm2.weight") or k == "norm.weight":
            P[k] = 1 + (torch.randn(shp, generator=g) * .1).clamp_(-.2, .2)
        elif len(shp) == 1:
            P[k] = (torch.randn(shp, generator=g) * .02).clamp_(-.04, .04)
        elif k in ("cls_token", "pos_embed"):
            P[k] = (torch.randn(shp, generator=g) * .02).clamp_(-.04, .04)
        else:
            std = .05 if k in ("patch_embed.proj.weight", "head.weight") else .04
            P[k] = (torch.randn(shp, generator=g) * std).clamp_(-2 * std, 2 * std)
    return P


def make_ft_inputs(cfg: SubnetCfg, batch: int, seed: int = 1, drop_path_rate: float = 0.0, soft: bool = False):
    """images, labels (or Mixup-like soft targets [B, C]), DropPath multipliers [depth, 2, B] (ones in eval-mode finetuning,
    finetune.py:445 / SURVEY App. B-10)."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(batch, 3, cfg.img, cfg.img, generator=g)
    labels = torch.randint(0, cfg.num_classes, (batch,), generator=g)
    scale = torch.ones(cfg.depth, 2, batch)
    dpr = torch.linspace(0, drop_path_rate, cfg.depth)
    for l in range(cfg.depth):
        p = float(dpr[l])
        if p > 0:
            u = torch.rand(2, batch, generator=g)
            scale[l] = torch.floor(1 - p + u) / (1 - p)
    target = None
    if soft:
        # what timm Mixup produces: lam * smoothed one-hot(y) + (1 - lam) * smoothed one-hot(y flipped)
        lam, sm, C = 0.7, cfg.smoothing, cfg.num_classes
        oh = lambda y: torch.full((batch, C), sm / C).scatter_(1, y.unsqueeze(1), 1 - sm + sm / C)
        target = lam * oh(labels) + (1 - lam) * oh(labels.flip(0))
    return images, labels, scale, target


def ft_forward(P: Dict[str, torch.Tensor], images: torch.Tensor, cfg: SubnetCfg, drop_scale: Optional[torch.Tensor] = None):
    """logits [B, C] (vision_transformer.py:332-358)."""
    B = images.shape[0]
    D = cfg.embed_dim
    x = F.conv2d(images, P["patch_embed.proj.weight"], P["patch_embed.proj.bias"], stride=cfg.patch)    # layers.py:126
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat([P["cls_token"].expand(B, -1, -1), x], dim=1) + P["pos_embed"]                          # vt:333-339
    N = x.shape[1]
    for l in range(cfg.depth):
        p, H, d = f"blocks.{l}.", cfg.heads[l], cfg.head_dims[l]
        s1 = drop_scale[l, 0].view(B, 1, 1) if drop_scale is not None else 1.0
        s2 = drop_scale[l, 1].view(B, 1, 1) if drop_scale is not None else 1.0
        y = F.layer_norm(x, (D,), P[p + "norm1.weight"], P[p + "norm1.bias"], cfg.eps)
        qkv = F.linear(y, P[p + "attn.qkv.weight"], P[p + "attn.qkv.bias"]).reshape(B, N, 3, H, d).permute(2, 0, 3, 1, 4)
        attn = ((qkv[0] @ qkv[1].transpose(-2, -1)) * cfg.scale).softmax(dim=-1)                          # layers.py:387-389
        o = (attn @ qkv[2]).transpose(1, 2).reshape(B, N, H * d)
        x = x + s1 * F.linear(o, P[p + "attn.proj.weight"], P[p + "attn.proj.bias"])                      # vt:158
        y = F.layer_norm(x, (D,), P[p + "norm2.weight"], P[p + "norm2.bias"], cfg.eps)
        h = F.gelu(F.linear(y, P[p + "mlp.fc1.weight"], P[p + "mlp.fc1.bias"]))
        x = x + s2 * F.linear(h, P[p + "mlp.fc2.weight"], P[p + "mlp.fc2.bias"])                          # vt:159
    x = F.layer_norm(x, (D,), P["norm.weight"], P["norm.bias"], cfg.eps)
    return F.linear(x[:, 0], P["head.weight"], P["head.bias"])


def ft_loss(logits, labels=None, target=None, smoothing=0.1):
    logp = F.log_softmax(logits, dim=-1)
    if target is not None:                                       # SoftTargetCrossEntropy
        return torch.sum(-target * logp, dim=-1).mean()
    nll = -logp.gather(-1, labels.unsqueeze(1)).squeeze(1)       # LabelSmoothingCrossEntropy
    return ((1 - smoothing) * nll + smoothing * (-logp.mean(-1))).mean()


def layer_id(name: str, depth: int) -> int:
    """lr_decay.get_layer_id_for_vit (lr_decay.py:62-75) with num_layers = depth + 1."""
    if name in ("cls_token", "pos_embed") or name.startswith("patch_embed"):
        return 0
    if name.startswith("blocks"):
        return int(name.split(".")[1]) + 1
    return depth + 1


def ft_group(name: str, shape, depth: int, weight_decay: float, layer_decay: float):
    """(group key, lr scale, weight decay) of a parameter: lr_decay.param_groups_lrd (lr_decay.py:15-59) with
    no_weight_decay_list = VisionTransformer.no_weight_decay() = pos_embed, cls_token, dist_token (vt:316-319)."""
    lid = layer_id(name, depth)
    nd = len(shape) == 1 or name in ("pos_embed", "cls_token", "dist_token")
    return (lid, 0 if nd else 1), layer_decay ** (depth + 1 - lid), 0.0 if nd else weight_decay


def torch_adamw_step(p, g, m, v, t, lr, wd, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.AdamW single-tensor update (decoupled decay first, bias-corrected Adam)."""
    p.mul_(1 - lr * wd)
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    denom = (v.sqrt() / math.sqrt(1 - b2 ** t)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / (1 - b1 ** t))


def ft_train_step(P, state, images, labels, cfg: SubnetCfg, lr: float, step: int, drop_scale=None, target=None,
                  weight_decay=0.05, layer_decay=0.95, update=True):
    """One train_one_epoch iteration (engine.py:31-62). Returns (logits, loss, grads); P / state updated in place."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    logits = ft_forward(leaves, images, cfg, drop_scale)
    loss = ft_loss(logits, labels, target, cfg.smoothing)
    loss.backward()
    grads = {k: leaves[k].grad.detach() for k in leaves}
    if update:
        for k in P:
            _, sc, wd = ft_group(k, P[k].shape, cfg.depth, weight_decay, layer_decay)
            st = state.setdefault(k, dict(m=torch.zeros_like(P[k]), v=torch.zeros_like(P[k])))
            torch_adamw_step(P[k], grads[k], st["m"], st["v"], step, lr * sc, wd)
    return logits.detach(), loss.detach(), grads
