"""Module-surface drop-in: the reference's own model object and training loop on top of the sm_100a engine.

The reference has no FFI: its seam is the `nn.Module` surface plus the `ModuleInjection` static factory (models/layers.py:
1052-1081) that `MAEBlock.__init__` / `MIMVisionTransformer.__init__` call (models/vision_transformer.py:181, 187, 434).
`install(layers_module, vt_module)` patches that factory so that a model built by the reference's constructor carries
  OFBPatchEmbed(MAEPatchEmbed) / OFBSparseAttention(MAESparseAttention) / OFBSparseMlp(MAESparseMlp)
- subclasses of the reference's classes, so the whole attribute / method contract (alpha, score, switch_cell, mask,
weighted_mask, weighted_mask_embed, w_p, finish_search, execute_prune, fused, head_num, ratio lists, update_w(), compress(),
compress_patchembed(), fuse(), get_alpha(), get_weight(), get_flops(), get_params_count(), pickling) is the reference's own
code - and replaces `MIMVisionTransformer.forward` with ONE whole-model autograd.Function over the engine:

    outputs, (decoder_loss, score_loss) = model(samples)          # engine.py:131 - unchanged caller
    forward : parameters -> engine arena (only tensors whose version changed), gates / LN / GEMMs / attention / decoder in the
              engine's kernels (SearchStepEngine.forward), returns logits [B, C] fp32 and the PMIM decoder loss (0-dim);
              every searchable module's `weighted_mask` (graph to alpha) and `weighted_mask_embed` are refreshed the way the
              reference's forward leaves them, because OFBSearchLOSS -> get_flops() reads them afterwards (losses.py:93-94).
    backward: takes d logits and d decoder_loss from the reference's criterion / loss weighting (engine.py:134-169), runs
              SearchStepEngine.backward(loss_grads=False) and hands every parameter its gradient; the sparsity / FLOPs loss
              gradients reach alpha / score through ordinary autograd (they are tiny torch expressions on the parameters).
The reference's optimizers, lr schedulers, compress() and checkpointing keep working on the model's nn.Parameters; after a
compress() that changes shapes the bridge rebuilds its engine from the model's current state (pruned widths, switch cells).
Random draws are made in the reference's order (PMIM noise, then two DropPath draws per block with a non-zero rate), so a run
is comparable step by step with a plain reference run from the same seed.

No fallback: a model that is not on a CUDA device, or a state the engine does not cover (fused model), raises OfbError.
Nothing here imports the reference - install() receives its modules from the caller.
"""
import math
from typing import Dict, List

import torch

from ._lib import OfbError
from .engine import (SearchStepEngine, embed_widths, head_channel_widths, head_counts, hidden_widths)

_STATE = {}


# ---------------------------------------------------------------------------------------------------------------------
# searchable-module subclasses
# ---------------------------------------------------------------------------------------------------------------------
def _alive_softmax(alpha, switch):
    a = alpha - torch.where(switch.to(alpha.device), torch.zeros_like(alpha), torch.full_like(alpha, float("inf")))
    return torch.softmax(a.view(-1), dim=0).reshape_as(alpha)


def make_classes(layers):
    """Subclasses of the reference's searchable modules (created against the caller's `models.layers` module)."""

    def _no_module_forward(self, *a, **k):
        raise OfbError(f"{type(self).__name__}: the B200 path runs fused at model level (MIMVisionTransformer.forward -> "
                       "SearchStepEngine); there is no per-module / eager fallback")

    class OFBPatchEmbed(layers.MAEPatchEmbed):
        ofb_kind = 0

        def refresh_weighted_mask(self):
            """self.weighted_mask exactly as MAEPatchEmbed.forward leaves it (layers.py:179-184): [1, D], graph to alpha."""
            if not self.finish_search:
                p = _alive_softmax(self.alpha, self.switch_cell) * self.switch_cell.to(self.alpha.device)
                self.weighted_mask = (p.view(-1, 1) * self.mask.to(p.device)).sum(0).unsqueeze(-2)

        forward = _no_module_forward

    class OFBSparseAttention(layers.MAESparseAttention):
        ofb_kind = 2

        def refresh_weighted_mask(self):
            """layers.py:494-496: [H, 1, d], sum over alive cells of softmax(alpha)_ij * mask[i, :, j, :]."""
            if not self.finish_search:
                p = _alive_softmax(self.alpha, self.switch_cell) * self.switch_cell.to(self.alpha.device)
                self.weighted_mask = torch.einsum("ij,ihjd->hd", p, self.mask.to(p.device)).unsqueeze(-2)

        forward = _no_module_forward

    class OFBSparseMlp(layers.MAESparseMlp):
        ofb_kind = 1

        def refresh_weighted_mask(self):
            """layers.py:847-852: [1, hidden]."""
            if not self.finish_search:
                p = _alive_softmax(self.alpha, self.switch_cell) * self.switch_cell.to(self.alpha.device)
                self.weighted_mask = (p.view(-1, 1) * self.mask.to(p.device)).sum(0).unsqueeze(-2)

        forward = _no_module_forward

    # whole-object pickles (torch.save(model), search.py:671-740) look the classes up by module path: publish them here
    # (install() must have run in the process that unpickles, exactly like the reference's own classes must be importable)
    for cls in (OFBPatchEmbed, OFBSparseAttention, OFBSparseMlp):
        cls.__module__ = __name__
        cls.__qualname__ = cls.__name__
        globals()[cls.__name__] = cls
    return OFBPatchEmbed, OFBSparseAttention, OFBSparseMlp


# ---------------------------------------------------------------------------------------------------------------------
# model <-> engine bridge
# ---------------------------------------------------------------------------------------------------------------------
class _FusedStep(torch.autograd.Function):
    """(logits, decoder_loss) = f(images, *parameters) on the engine; backward returns the engine's gradients."""

    @staticmethod
    def forward(ctx, bridge, images, noise, drop_u, *params):
        eng = bridge.engine
        eng.forward(images, bridge.dummy_labels, noise=noise, drop_u=drop_u)
        ctx.bridge = bridge
        ctx.pmim = eng._pmim
        logits = eng.logits.clone()
        dec = eng.scal[2].clone() if eng._pmim else torch.zeros((), device=logits.device)
        return logits, dec

    @staticmethod
    def backward(ctx, dlogits, ddec):
        bridge = ctx.bridge
        eng = bridge.engine
        eng.dlogits.copy_(dlogits)                              # fp32 -> bf16: the head data / weight gradient GEMM operand
        if ctx.pmim:
            # decoder_loss = sum|.| / ((n_masked * 256 + 1e-5) * 3): the backward GEMMs take d loss / d x_rec per unit sign
            eng.scal[5:6] = ddec.reshape(1).to(torch.float32) / ((eng.scal[6:7] * 256.0 + 1e-5) * 3.0)
        eng.grads.zero_()
        eng.backward(loss_grads=False)
        grads = tuple(eng._unpad(k, eng.g(k)).clone().reshape(shape) for k, shape in zip(bridge.names, bridge.param_shapes))
        return (None, None, None, None) + grads


class ModelBridge:
    """One per model instance: owns the SearchStepEngine that mirrors the model's current shapes."""

    # class-level defaults: a bridge that travelled through pickle (whole-object checkpoints, search.py:671-740) comes back
    # empty and rebuilds its engine on the next forward
    engine = None
    signature = None
    dummy_labels = None

    def __init__(self, model=None):
        self.versions = {}
        self.names: List[str] = []
        self.param_shapes = []

    def __getstate__(self):
        return {"versions": {}, "names": [], "param_shapes": []}       # never the engine (device arenas, ctypes handles)

    # ---- model state -> engine configuration ----
    @staticmethod
    def searchables(model):
        mods = {"patch_embed": model.patch_embed}
        for l, blk in enumerate(model.blocks):
            mods[f"blocks.{l}.attn"] = blk.attn
            mods[f"blocks.{l}.mlp"] = blk.mlp
        return mods

    def _config(self, model, batch):
        mods = self.searchables(model)
        for k, m in mods.items():
            if not hasattr(m, "refresh_weighted_mask"):
                raise OfbError(f"{k} is a {type(m).__name__}: build the model after ofb_b200.modules.install() so that "
                               "ModuleInjection creates the OFB* modules")
            if getattr(m, "fused", False):
                raise OfbError("fused model: hand the folded weights to FinetuneStepEngine (prune.fuse_params / subnet_dims)")
        blk0 = model.blocks[0]
        D0, H0, d0 = model.embed_dim, blk0.attn.num_heads, blk0.attn.head_dim
        hid0 = blk0.mlp.hidden_features
        depth = len(model.blocks)
        Dv = model.pos_embed.shape[-1]
        heads = [int(b.attn.score.shape[0]) for b in model.blocks]
        hdims = [int(b.attn.score.shape[1]) for b in model.blocks]
        hids = [int(b.mlp.score.shape[-1]) for b in model.blocks]
        switches, spaces = {}, {}
        ew, hc, cw, hw = embed_widths(D0), head_counts(H0), head_channel_widths(d0), hidden_widths(hid0)
        for k, m in mods.items():
            sw = m.switch_cell.detach().to("cpu", torch.bool)
            if m.ofb_kind == 2:
                sw = sw.reshape(-1, sw.shape[-1])
                spaces[k] = (cw[:sw.shape[1]], hc[:sw.shape[0]])
            else:
                sw = sw.reshape(1, -1)
                spaces[k] = ((ew if m.ofb_kind == 0 else hw)[:sw.shape[1]], [])
            switches[k] = sw
        full = (Dv == D0 and heads == [H0] * depth and hdims == [d0] * depth and hids == [hid0] * depth
                and all(spaces[k][0] == (ew if m.ofb_kind == 0 else cw if m.ofb_kind == 2 else hw) for k, m in mods.items())
                and all(spaces[k][1] == hc for k, m in mods.items() if m.ofb_kind == 2))
        pruned = None if full else dict(embed=Dv, heads=heads, head_dims=hdims, hiddens=hids, spaces=spaces)
        dpr = 0.0
        for b in model.blocks:
            dpr = max(dpr, float(getattr(b.drop_path, "drop_prob", 0.0) or 0.0))
        sig = (batch, D0, H0, depth, hid0, Dv, tuple(heads), tuple(hdims), tuple(hids), dpr, model.num_classes,
               tuple((k, tuple(v.shape), bytes(v.reshape(-1).to(torch.uint8).tolist())) for k, v in switches.items()))
        return sig, dict(embed_dim=D0, num_heads=H0, depth=depth, batch=batch, mlp_ratio=hid0 // D0, num_classes=model.num_classes,
                         drop_path_rate=dpr, switches=switches, pruned=pruned)

    def _ensure_engine(self, model, images):
        sig, cfg = self._config(model, images.shape[0])
        if sig != self.signature:
            self.engine = None                      # release the old arenas before the new ones are allocated
            self.engine = SearchStepEngine(device=images.device, **cfg)
            self.signature = sig
            self.versions = {}
            self.dummy_labels = torch.zeros(images.shape[0], dtype=torch.int64, device=images.device)
        return self.engine

    def _sync_params(self, model):
        """Copy every parameter whose storage or version changed since the last step into the engine arena (the reference's
        optimizers update the nn.Parameters in place; compress() re-creates them)."""
        eng = self.engine
        named = dict(model.named_parameters())
        missing = [k for k in eng.offsets if k not in named]
        if missing:
            raise OfbError(f"model lacks parameters the engine expects: {missing[:4]}")
        dirty = False
        with torch.no_grad():
            for k in eng.offsets:
                p = named[k]
                tag = (p.data_ptr(), p._version)
                if self.versions.get(k) != tag:
                    eng.p(k).copy_(eng._pad(k, p.detach().to(eng.dev, torch.float32).reshape(eng.ref_shapes[k])))
                    self.versions[k] = tag
                    dirty = True
        if dirty:
            eng.sync_shadow()
        self.names = list(eng.offsets)
        self.param_shapes = [tuple(named[k].shape) for k in self.names]
        return [named[k] for k in self.names]

    # ---- the model's forward ----
    def forward(self, model, images):
        if not images.is_cuda:
            raise OfbError("ofb_b200 modules need CUDA tensors (no CPU fallback)")
        eng = self._ensure_engine(model, images)
        params = self._sync_params(model)
        mods = self.searchables(model)
        # schedule state the reference keeps on the modules / the model (update_w, adjust_masking_ratio; engine.py:104-117)
        w_ps = {float(m.w_p) for m in mods.values() if not m.finish_search}
        if len(w_ps) > 1:
            raise OfbError(f"searchable modules disagree on w_p: {sorted(w_ps)}")
        eng.w_p = w_ps.pop() if w_ps else eng.w_p
        keeps = [r for i, r in enumerate(model.patch_ratio_list) if bool(model.switch_cell_patch[:, i])]
        eng.keep_ratio = float(keeps[0]) if len(keeps) == 1 else 1.0
        eng._schedule_touched = True
        eng._hyper_up.begin(keep=True)[eng._wp_idx] = eng.w_p
        eng._hyper_up.upload(eng.hyper)
        # what OFBSearchLOSS / get_flops() read after the forward (graph to the alphas)
        for m in mods.values():
            m.refresh_weighted_mask()
        pe = model.patch_embed
        wme = pe.get_weight()[0] if not pe.finish_search else pe.weighted_mask
        for blk in model.blocks:
            blk.weighted_mask_embed = blk.attn.weighted_mask_embed = blk.mlp.weighted_mask_embed = wme
        if not model.training:
            # evaluate() (engine.py:222-257): gates as in training, no PMIM, identity DropPath, no decoder branch
            with torch.no_grad():
                eng.forward(images, self.dummy_labels, train=False)
                return eng.logits.clone(), (0., None)
        # random draws in the reference's order: PMIM noise (vt:597), then DropPath of attn / mlp per block (timm DropPath)
        B, L = images.shape[0], eng.L
        keep = int(L * eng.keep_ratio)
        noise = torch.rand(B, L, device=images.device) if keep != L else None
        drop_u = torch.ones(eng.depth * 2, B, device=images.device)
        for l, blk in enumerate(model.blocks):
            if float(getattr(blk.drop_path, "drop_prob", 0.0) or 0.0) > 0.0:
                drop_u[2 * l] = torch.rand((B, 1, 1), dtype=images.dtype, device=images.device).view(B)
                drop_u[2 * l + 1] = torch.rand((B, 1, 1), dtype=images.dtype, device=images.device).view(B)
        # per-block drop probabilities as the model holds them (linspace(0, rate, depth), vt:447)
        eng.drop_prob.copy_(torch.tensor([float(getattr(b.drop_path, "drop_prob", 0.0) or 0.0) for b in model.blocks],
                                         device=eng.drop_prob.device).repeat_interleave(2))
        logits, dec = _FusedStep.apply(self, images, noise, drop_u, *params)
        decoder_loss = dec if eng._pmim else 0.
        return logits, (decoder_loss, None)


def _fused_forward(self, imgs):
    if not hasattr(self.patch_embed, "refresh_weighted_mask"):
        # a model that was built WITHOUT the patched factory (plain reference modules) keeps the reference's own forward
        return _STATE["orig"][3](self, imgs)
    bridge = self.__dict__.get("_ofb_bridge")
    if bridge is None:
        bridge = ModelBridge(self)
        object.__setattr__(self, "_ofb_bridge", bridge)
    return bridge.forward(self, imgs)


def install(layers, vt):
    """Patch the reference's seam: `layers` = its models.layers module (ModuleInjection, MAE* classes), `vt` = its
    models.vision_transformer module (MIMVisionTransformer). Idempotent; uninstall() restores the originals."""
    if _STATE.get("installed"):
        return _STATE["classes"]
    PE, AT, ML = make_classes(layers)
    MI = layers.ModuleInjection
    _STATE.update(installed=True, classes=(PE, AT, ML), layers=layers, vt=vt,
                  orig=(MI.__dict__["make_searchable_patchembed"], MI.__dict__["make_searchable_maeattn"],
                        MI.__dict__["make_searchable_maemlp"], vt.MIMVisionTransformer.forward))

    def make_searchable_patchembed(patchmodule, embed_search=True):
        if MI.method == "full":
            return patchmodule
        m = PE(patchmodule, embed_search)
        if embed_search:
            MI.searchable_modules.append(m)
        return m

    def make_searchable_maeattn(attn_module, head_search=False, channel_search=False, attn_search=True):
        if MI.method == "full":
            return attn_module
        m = AT(attn_module, head_search, channel_search, attn_search)
        if attn_search:
            MI.searchable_modules.append(m)
        return m

    def make_searchable_maemlp(mlp_module, mlp_search=True):
        if MI.method == "full":
            return mlp_module
        m = ML(mlp_module, mlp_search)
        if mlp_search:
            MI.searchable_modules.append(m)
        return m

    MI.make_searchable_patchembed = staticmethod(make_searchable_patchembed)
    MI.make_searchable_maeattn = staticmethod(make_searchable_maeattn)
    MI.make_searchable_maemlp = staticmethod(make_searchable_maemlp)
    vt.MIMVisionTransformer.forward = _fused_forward
    return PE, AT, ML


def uninstall():
    if not _STATE.get("installed"):
        return
    MI = _STATE["layers"].ModuleInjection
    o = _STATE["orig"]
    MI.make_searchable_patchembed, MI.make_searchable_maeattn, MI.make_searchable_maemlp = o[0], o[1], o[2]
    _STATE["vt"].MIMVisionTransformer.forward = o[3]
    _STATE.clear()


def strip(model):
    """Drop the bridge (engine arenas) from a model, e.g. to free its memory before evaluation with another batch size.
    Pickling the whole model (search.py:671-740) works without it: ModelBridge.__getstate__ leaves the engine out."""
    model.__dict__.pop("_ofb_bridge", None)
    return model
