"""Mixup / CutMix of the post-search phase and the finetune loop (search.py:651-655 `Mixup(mixup_alpha=0.8, cutmix_alpha=1.0,
prob=args.mixup_prob, switch_prob=args.mixup_switch_prob, mode='batch', label_smoothing=args.smoothing)`; applied at
engine.py:98-99 and finetune.py:360-366).

The reference takes this from timm (timm.data.Mixup, third-party). As there, the per-batch parameters - lam, mixup vs cutmix, the
box - are drawn on the HOST from numpy's global RNG in timm's call order (Mixup._params_per_batch, rand_bbox,
cutmix_bbox_and_lam with correct_lam=True), so a run seeded like the reference (np.random.seed, search.py:383) makes the same
choices; the arithmetic on the batch runs on the device (ofb_mixup_batch / ofb_patchify_mixup / ofb_mixup_target)."""
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from . import ops


@dataclass
class MixParams:
    lam: float = 1.0
    box: Optional[Tuple[int, int, int, int]] = None     # CutMix (yl, yh, xl, xh); None = Mixup blend

    @property
    def identity(self):
        return self.lam == 1.0


class Mixup:
    """mode='batch' only (the reference's default, search.py:155); cutmix_minmax is not supported (default None)."""

    def __init__(self, mixup_alpha=0.8, cutmix_alpha=1.0, prob=1.0, switch_prob=0.5, label_smoothing=0.1, num_classes=1000,
                 rng=None):
        self.mixup_alpha, self.cutmix_alpha, self.prob, self.switch_prob = mixup_alpha, cutmix_alpha, prob, switch_prob
        self.label_smoothing, self.num_classes = label_smoothing, num_classes
        self.rng = rng if rng is not None else np.random          # timm draws from the global numpy RNG

    def draw(self, img_hw=(224, 224)) -> MixParams:
        rng = self.rng
        lam, use_cutmix = 1.0, False
        if rng.rand() < self.prob:
            if self.mixup_alpha > 0. and self.cutmix_alpha > 0.:
                use_cutmix = rng.rand() < self.switch_prob
                lam = rng.beta(self.cutmix_alpha, self.cutmix_alpha) if use_cutmix else rng.beta(self.mixup_alpha, self.mixup_alpha)
            elif self.mixup_alpha > 0.:
                lam = rng.beta(self.mixup_alpha, self.mixup_alpha)
            elif self.cutmix_alpha > 0.:
                use_cutmix, lam = True, rng.beta(self.cutmix_alpha, self.cutmix_alpha)
            lam = float(lam)
        if lam == 1.0 or not use_cutmix:
            return MixParams(lam, None)
        H, W = img_hw
        ratio = np.sqrt(1 - lam)
        cut_h, cut_w = int(H * ratio), int(W * ratio)
        cy, cx = rng.randint(0, H), rng.randint(0, W)
        yl, yh = int(np.clip(cy - cut_h // 2, 0, H)), int(np.clip(cy + cut_h // 2, 0, H))
        xl, xh = int(np.clip(cx - cut_w // 2, 0, W)), int(np.clip(cx + cut_w // 2, 0, W))
        return MixParams(1. - (yh - yl) * (xh - xl) / float(H * W), (yl, yh, xl, xh))

    def __call__(self, images, labels, target_out, params: Optional[MixParams] = None, images_out=None):
        """timm's `x, target = mixup_fn(x, target)`: mixes `images` (in place unless images_out is given) and fills the soft
        targets `target_out` fp32 [B, C]. Returns the parameters used."""
        assert images.shape[0] % 2 == 0, "Batch size should be even when using this"          # timm's own assert
        mp = params if params is not None else self.draw(tuple(images.shape[-2:]))
        if not mp.identity:
            ops.mixup_batch(images, images if images_out is None else images_out, mp.lam, mp.box)
        elif images_out is not None:
            images_out.copy_(images)
        ops.mixup_target(labels, target_out, mp.lam, self.label_smoothing)
        return mp
