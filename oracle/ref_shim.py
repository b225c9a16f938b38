"""TEST INFRASTRUCTURE ONLY — import shim that lets the *unmodified* reference under /root/reference be imported in
this container (no timm / apex / matplotlib installed, torch >= 2 has no torch._six).

Used by oracle/make_golden*.py to pin the oracle restatement (oracle/ofb_oracle.py) against the real reference and to
produce tests/golden/*.npz, and by oracle/ref_runner.py (the reference arm of bench.py and the drop-in tests).
/root/reference does not exist on the GPU box: there the byte-identical staged copy oracle/_ref/ (oracle/make_ref.py,
git-ignored, shipped with the snapshot) is imported instead.

The stand-ins restate the standard timm-0.4 semantics of the handful of third-party symbols the hot path touches
(SURVEY.md §8c): trunc_normal_, to_2tuple, DropPath, LabelSmoothingCrossEntropy, SoftTargetCrossEntropy, accuracy.
"""
import math
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

import os

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = "/root/reference" if os.path.isdir("/root/reference") else _STAGED


def available() -> bool:
    """The unmodified reference can be imported here (build container: /root/reference; GPU box: the staged oracle/_ref)."""
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "layers.py"))


def _trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    # timm/models/layers/weight_init.py semantics == torch.nn.init.trunc_normal_
    return torch.nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


def _lecun_normal_(tensor):
    fan_in = tensor.shape[1] if tensor.dim() > 1 else tensor.shape[0]
    return _trunc_normal_(tensor, std=math.sqrt(1.0 / fan_in) / .87962566103423978)


def _to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class DropPath(nn.Module):
    """timm DropPath: x / keep * floor(keep + U[0,1)) with one draw per sample."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0. or not self.training:
            return x
        keep = 1 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        rt = keep + torch.rand(shape, dtype=x.dtype, device=x.device)
        rt.floor_()
        return x.div(keep) * rt


class LabelSmoothingCrossEntropy(nn.Module):
    def __init__(self, smoothing=0.1):
        super().__init__()
        self.smoothing = smoothing
        self.confidence = 1. - smoothing

    def forward(self, x, target):
        logprobs = F.log_softmax(x, dim=-1)
        nll = -logprobs.gather(dim=-1, index=target.unsqueeze(1)).squeeze(1)
        smooth = -logprobs.mean(dim=-1)
        return (self.confidence * nll + self.smoothing * smooth).mean()


class SoftTargetCrossEntropy(nn.Module):
    def forward(self, x, target):
        return torch.sum(-target * F.log_softmax(x, dim=-1), dim=-1).mean()


def _accuracy(output, target, topk=(1,)):
    maxk = max(topk)
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.t().eq(target.reshape(1, -1).expand_as(pred.t()))
    return [correct[:k].reshape(-1).float().sum(0) * 100. / target.size(0) for k in topk]


def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install():
    """Register the stand-in modules and put the reference on sys.path. Idempotent."""
    if "timm" in sys.modules and getattr(sys.modules["timm"], "_ofb_shim", False):
        return
    ident = lambda f=None, **kw: f if f is not None else (lambda g: g)
    timm = _mod("timm", _ofb_shim=True)
    _mod("timm.models")
    layers = _mod("timm.models.layers", trunc_normal_=_trunc_normal_, lecun_normal_=_lecun_normal_, DropPath=DropPath,
                  to_2tuple=_to_2tuple)
    _mod("timm.models.layers.helpers", to_2tuple=_to_2tuple)
    _mod("timm.models.helpers", build_model_with_cfg=None, overlay_external_default_cfg=None)
    _mod("timm.models.registry", register_model=ident)
    _mod("timm.models.vision_transformer", VisionTransformer=object, _cfg=lambda **kw: kw)
    _mod("timm.data", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406), IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225),
         IMAGENET_INCEPTION_MEAN=(0.5, 0.5, 0.5), IMAGENET_INCEPTION_STD=(0.5, 0.5, 0.5), Mixup=object,
         create_transform=None)
    _mod("timm.utils", accuracy=_accuracy, ModelEma=object, NativeScaler=object, get_state_dict=None)
    _mod("timm.loss", LabelSmoothingCrossEntropy=LabelSmoothingCrossEntropy,
         SoftTargetCrossEntropy=SoftTargetCrossEntropy)
    _mod("timm.scheduler")
    _mod("timm.scheduler.scheduler", Scheduler=object)
    _mod("timm.scheduler.cosine_lr", CosineLRScheduler=object)
    _mod("timm.optim", create_optimizer=None)
    timm.models = sys.modules["timm.models"]
    sys.modules["timm.models"].layers = layers
    _mod("matplotlib")
    _mod("matplotlib.pyplot")
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    _mod("apex", amp=types.SimpleNamespace())
    _mod("apex.amp")
    if not hasattr(torch, "_six"):
        _mod("torch._six", inf=math.inf)
    if not torch.cuda.is_available():
        torch.cuda.synchronize = lambda *a, **k: None
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def build_reference_model(embed_dim, num_heads, depth=12, drop_path_rate=0.1, num_classes=1000, seed=0):
    """MIMVisionTransformer exactly as search.py builds it for `--attn_search --mlp_search --embed_search --mae`
    (SURVEY.md §8d), pretrained=False."""
    install()
    from functools import partial
    from models.layers import LayerNorm, ModuleInjection, PatchEmbed
    from models.vision_transformer import MIMVisionTransformer
    torch.manual_seed(seed)
    ModuleInjection.method = "search"
    ModuleInjection.searchable_modules = []
    model = MIMVisionTransformer(
        patch_size=16, embed_dim=embed_dim, depth=depth, num_heads=num_heads, mlp_ratio=4, qkv_bias=True,
        norm_layer=partial(LayerNorm, eps=1e-6), embed_layer=PatchEmbed, mae=True, attn_search=True, mlp_search=True,
        embed_search=True, patch_search=False, mask_ratio=1.0, norm_pix_loss=True, drop_path_rate=drop_path_rate,
        num_classes=num_classes)
    model.searchable_modules = [m for m in model.modules() if hasattr(m, "alpha")]
    model.correct_require_grad(0.5, 0.5, 0, 0.5)
    return model


class FakeDDP(nn.Module):
    """The reference dereferences model.module unconditionally (losses.py:93-94, engine.py:204)."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)
