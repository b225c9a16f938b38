"""FinetuneStepEngine — one *finetune* training step of a physically pruned (searched) subnet on the same sm_100a kernels
as the search step (BASELINE.json configs[4]).

Reference step: engine.train_one_epoch loop body (engine.py:31-62) = plain VisionTransformer.forward
(vision_transformer.py:332-358; pre-norm Block 157-160; Attention layers.py:382-394; Mlp 784-790) on the per-layer shapes
finetune.intersect installs (finetune.py:182-249) -> LabelSmoothingCrossEntropy / SoftTargetCrossEntropy
(finetune.py:388-394) -> backward -> torch.optim.AdamW over lr_decay.param_groups_lrd (lr_decay.py:15-75,
finetune.py:378-383) [-> DDP all-reduce].

Pruned shapes and the hardware: embedding widths after pruning are multiples of 12 (6 for DeiT-T), head dims multiples of 8.
  * the embedding axis is stored padded to a multiple of 8 channels (16-byte rows for TMA); the <= 7 padding channels hold
    zeros everywhere (activations, weights, LayerNorm gamma/beta), LayerNorm normalises over the real channels only
    (ofb_layernorm_*_ex) and returns zero gradient for the padding, so the padding stays exactly zero under AdamW;
  * every head is stored 64 wide (the attention kernel's head dim): the 64 - d'_l extra channels of q, k, v are zero rows of
    the qkv weight / zero columns of the proj weight - they add nothing to q k^T or to P v and receive zero gradient. The
    softmax scale is the constant (D/H)^-0.5 of the unpruned model, exactly as in the reference;
  * head counts and MLP hidden widths are used as they are (the GEMMs take any M / N / K that is a multiple of 8).
Parameters are exchanged in reference shapes (load_params / named_parameters / named_grads pad and unpad).
There is no CPU path: every op goes through the C ABI in libofb_b200.so.
"""
import ctypes as C
import math
import os
from typing import Dict, List, Optional

import torch

from . import dp, ops

T_PAD = 8
HD = 64          # physical head width of the attention kernel


def layer_id(name: str, depth: int) -> int:
    """lr_decay.get_layer_id_for_vit (lr_decay.py:62-75)."""
    if name in ("cls_token", "pos_embed") or name.startswith("patch_embed"):
        return 0
    if name.startswith("blocks"):
        return int(name.split(".")[1]) + 1
    return depth + 1


def is_no_decay(name: str, ref_shape) -> bool:
    """lr_decay.param_groups_lrd (lr_decay.py:31-37) with VisionTransformer.no_weight_decay() (vt:316-319)."""
    return len(ref_shape) == 1 or name in ("pos_embed", "cls_token", "dist_token")


class FinetuneStepEngine:
    def __init__(self, embed_dim: int, heads: List[int], head_dims: List[int], hiddens: List[int], batch: int, *,
                 num_classes=1000, img=224, patch=16, attn_scale=0.125, lr=1e-3, weight_decay=0.05, layer_decay=0.95,
                 eps_ln=1e-6, smoothing=0.1, drop_path_rate=0.0, training_mode=False, accum_iter=1, device="cuda",
                 process_group=None):
        depth = len(heads)
        assert len(head_dims) == depth and len(hiddens) == depth
        assert all(0 < d <= HD and d % 8 == 0 for d in head_dims), "head dims: multiples of 8 up to 64"
        assert all(h % 8 == 0 for h in hiddens), "hidden widths: multiples of 8"
        self.Dv, self.Dp = embed_dim, (embed_dim + 7) // 8 * 8
        self.heads, self.head_dims, self.hiddens, self.depth, self.B = list(heads), list(head_dims), list(hiddens), depth, batch
        self.C, self.img, self.P = num_classes, img, patch
        self.L = (img // patch) ** 2
        self.T = self.L + 1
        self.M, self.ML = batch * self.T, batch * self.L
        self.dev = torch.device(device)
        self.scale = attn_scale
        self.lr, self.wd, self.layer_decay, self.eps_ln, self.smoothing = lr, weight_decay, layer_decay, eps_ln, smoothing
        self.accum_iter = accum_iter
        # finetune.py:445: a model finetuned from a checkpoint runs in eval mode -> DropPath is the identity
        self.training_mode, self.drop_path_rate = training_mode, drop_path_rate
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.step_count = 0

        # ---- parameter arenas: tensors ordered by (layer id, no-decay | decay) = the optimizer groups of lr_decay.py ----
        self.ref_shapes = self._ref_shapes()
        self.shapes = {k: self._padded_shape(k, s) for k, s in self.ref_shapes.items()}
        gkey = lambda k: (layer_id(k, depth), 0 if is_no_decay(k, self.ref_shapes[k]) else 1)
        order = sorted(self.ref_shapes, key=gkey)              # stable: keeps model order inside a group
        self.offsets, self.groups, seg_end, off = {}, [], [], 0
        for k in order:
            if not self.groups or self.groups[-1] != gkey(k):
                if self.groups:
                    seg_end.append(off)
                self.groups.append(gkey(k))
            self.offsets[k] = off
            off += (math.prod(self.shapes[k]) + T_PAD - 1) // T_PAD * T_PAD
        seg_end.append(off)
        self.n_arena, self.seg_end = off, seg_end
        self._seg_end_c = (C.c_int64 * len(seg_end))(*seg_end)
        f32 = dict(dtype=torch.float32, device=self.dev)
        bf = dict(dtype=torch.bfloat16, device=self.dev)
        self.params, self.grads = torch.zeros(off, **f32), torch.zeros(off, **f32)
        self.adam_m, self.adam_v = torch.zeros(off, **f32), torch.zeros(off, **f32)
        self.shadow = torch.zeros(off, **bf)
        nh = len(self.groups) * 8
        self._hyper_up = ops.HyperUploader(nh, self.dev)
        self.hyper = torch.zeros(nh, **f32)

        # ---- activations ----
        M, ML, Dp, B, T = self.M, self.ML, self.Dp, batch, self.T
        self.ldT = (M + 7) // 8 * 8
        self.mlp_bn = int(os.environ.get("OFB_MLP_BN", "256"))
        self.mlp_parts = ops.gemm_mlp_partial_rows(M, self.mlp_bn)
        self.patches = torch.empty(ML, 3 * patch * patch, **bf)
        self.zero_mask = torch.zeros(B, self.L, **f32)          # no PMIM in finetuning: every patch is kept
        self.ones = torch.ones(max(Dp, max(hiddens), max(heads) * HD), **f32)     # identity gate of the shared epilogues
        dpr = torch.linspace(0, drop_path_rate, depth).repeat_interleave(2)
        self.drop_prob = dpr.to(self.dev)
        self.drop_scale = torch.ones(depth * 2, B, **f32)
        self.xs = [torch.zeros(M, Dp, **bf) for _ in range(depth + 1)]
        self.blk = []
        for l in range(depth):
            A, hid = heads[l] * HD, hiddens[l]
            self.blk.append(dict(
                mean1=torch.empty(M, **f32), rstd1=torch.empty(M, **f32), x1=torch.empty(M, Dp, **bf),
                qkv=torch.empty(M, 3 * A, **bf), lse=torch.empty(B, heads[l], T, **f32), o=torch.empty(M, A, **bf),
                x2=torch.empty(M, Dp, **bf), mean2=torch.empty(M, **f32), rstd2=torch.empty(M, **f32),
                x3=torch.empty(M, Dp, **bf), u=torch.empty(hid, self.ldT, **bf), h=torch.empty(hid, self.ldT, **bf)))
        self.meanf, self.rstdf = torch.empty(M, **f32), torch.empty(M, **f32)
        self.latent = torch.empty(M, Dp, **bf)
        self.logits = torch.empty(B, self.C, **f32)
        self.loss_rows = torch.empty(B, **f32)
        self.dlogits = torch.empty(B, self.C, **bf)
        self.scal = torch.zeros(8, **f32)
        self.eval_rows, self.eval_out = torch.zeros(B, 3, **f32), torch.zeros(3, **f32)
        # backward scratch
        Amax, hmax = max(heads) * HD, max(hiddens)
        self.gA, self.gB, self.gC = (torch.zeros(M, Dp, **bf) for _ in range(3))
        self.dO = torch.empty(M, Amax, **bf)
        self.du = torch.empty(hmax, self.ldT, **bf)
        self.dqkv = torch.empty(M, 3 * Amax, **bf)
        self.dconv = torch.empty(ML, Dp, **bf)
        self.ln_parts = ops.layernorm_bwd_parts(M)
        self.pg_, self.pb_, self.pd_ = (torch.empty(self.ln_parts, Dp, **f32) for _ in range(3))
        self.pg2_, self.pb2_, self.pd2_ = (torch.empty(self.ln_parts, Dp, **f32) for _ in range(3))    # LayerNorm 2 (see backward)
        self.cp0, self.cp1 = torch.empty(self.mlp_parts, hmax, **f32), torch.empty(self.mlp_parts, hmax, **f32)
        self.att_pg, self.att_pb = torch.empty(B, Amax, **f32), torch.empty(B, 3 * Amax, **f32)
        self.e_gx, self.e_pos, self.e_mt = (torch.empty(T, Dp, **f32) for _ in range(3))
        # the exchange that follows backward is ONE all-reduce of the arena: on 8 x B200 89.6 MB take 0.30 ms in one piece, 0.46 ms
        # in four (profiles/r02b_allreduce_8gpu.txt); OFB_DP_BUCKETS restores a bucketed exchange
        self._dp_bounds = dp.bucket_bounds(self.n_arena, max_buckets=int(os.environ.get("OFB_DP_BUCKETS", "1")))
        self._graphs = {}

    # ------------------------------------------------------------------------------------------------------------
    def _ref_shapes(self):
        """state_dict names / shapes of the reference's finetune model, in named_parameters() order."""
        D, L, Cn = self.Dv, self.L, self.C
        s = {"cls_token": (1, 1, D), "pos_embed": (1, L + 1, D),
             "patch_embed.proj.weight": (D, 3, self.P, self.P), "patch_embed.proj.bias": (D,)}
        for l in range(self.depth):
            p, A, hid = f"blocks.{l}.", self.heads[l] * self.head_dims[l], self.hiddens[l]
            s[p + "norm1.weight"] = (D,); s[p + "norm1.bias"] = (D,)
            s[p + "attn.qkv.weight"] = (3 * A, D); s[p + "attn.qkv.bias"] = (3 * A,)
            s[p + "attn.proj.weight"] = (D, A); s[p + "attn.proj.bias"] = (D,)
            s[p + "norm2.weight"] = (D,); s[p + "norm2.bias"] = (D,)
            s[p + "mlp.fc1.weight"] = (hid, D); s[p + "mlp.fc1.bias"] = (hid,)
            s[p + "mlp.fc2.weight"] = (D, hid); s[p + "mlp.fc2.bias"] = (D,)
        s["norm.weight"] = (D,); s["norm.bias"] = (D,)
        s["head.weight"] = (Cn, D); s["head.bias"] = (Cn,)
        return s

    def _blk(self, name):
        return int(name.split(".")[1])

    def _padded_shape(self, name, shp):
        Dp = self.Dp
        if name in ("cls_token", "pos_embed"):
            return shp[:-1] + (Dp,)
        if name == "patch_embed.proj.weight":
            return (Dp, 3 * self.P * self.P)
        if name.endswith("attn.qkv.weight"):
            return (3 * self.heads[self._blk(name)] * HD, Dp)
        if name.endswith("attn.qkv.bias"):
            return (3 * self.heads[self._blk(name)] * HD,)
        if name.endswith("attn.proj.weight"):
            return (Dp, self.heads[self._blk(name)] * HD)
        if name.endswith("mlp.fc1.weight") or name == "head.weight":
            return (shp[0], Dp)
        if name.endswith("mlp.fc1.bias") or name == "head.bias":
            return shp
        if name.endswith("mlp.fc2.weight"):
            return (Dp, shp[1])
        return (Dp,)          # every remaining tensor is a vector over the embedding axis

    def _pad(self, name, t):
        """reference-shaped tensor -> zero-padded arena layout."""
        Dv, Dp = self.Dv, self.Dp
        out = torch.zeros(self.shapes[name], dtype=t.dtype, device=t.device)
        if "attn.qkv" in name or name.endswith("attn.proj.weight"):
            l = self._blk(name)
            H, d = self.heads[l], self.head_dims[l]
            if name.endswith("qkv.weight"):
                out.view(3, H, HD, Dp)[:, :, :d, :Dv] = t.reshape(3, H, d, Dv)
            elif name.endswith("qkv.bias"):
                out.view(3, H, HD)[:, :, :d] = t.reshape(3, H, d)
            else:
                out.view(Dp, H, HD)[:Dv, :, :d] = t.reshape(Dv, H, d)
            return out
        t2 = t.reshape(t.shape[0], -1) if name == "patch_embed.proj.weight" else t
        idx = tuple(slice(0, s) for s in t2.shape)
        out[idx] = t2
        return out

    def _unpad(self, name, t):
        """arena layout -> reference-shaped tensor."""
        Dv, Dp = self.Dv, self.Dp
        ref = self.ref_shapes[name]
        if "attn.qkv" in name or name.endswith("attn.proj.weight"):
            l = self._blk(name)
            H, d = self.heads[l], self.head_dims[l]
            if name.endswith("qkv.weight"):
                return t.view(3, H, HD, Dp)[:, :, :d, :Dv].reshape(ref)
            if name.endswith("qkv.bias"):
                return t.view(3, H, HD)[:, :, :d].reshape(ref)
            return t.view(Dp, H, HD)[:Dv, :, :d].reshape(ref)
        if name == "patch_embed.proj.weight":
            return t[:Dv].reshape(ref)
        return t[tuple(slice(0, s) for s in ref)].reshape(ref)

    def _view(self, arena, name):
        o, shp = self.offsets[name], self.shapes[name]
        return arena[o:o + math.prod(shp)].view(shp)

    def p(self, name):
        return self._view(self.params, name)

    def g(self, name):
        return self._view(self.grads, name)

    def w(self, name):
        v = self._view(self.shadow, name)
        return v.reshape(v.shape[0], -1)

    def load_params(self, named: Dict[str, torch.Tensor]):
        for k in self.offsets:
            self.p(k).copy_(self._pad(k, named[k].to(self.dev, torch.float32)))
        ops.cast_bf16(self.params, self.shadow)

    def init_params(self, seed=0):
        """Reference-style initialisation of the pruned shapes (trunc-normal .02 weights and tokens, zero biases, unit LayerNorm
        weights: vision_transformer.py:300-312); benchmarks only - a real run loads the searched weights."""
        g = torch.Generator(device="cpu").manual_seed(seed)
        named = {}
        for k, shp in self.ref_shapes.items():
            if k.endswith("norm1.weight") or k.endswith("norm2.weight") or k == "norm.weight":
                named[k] = torch.ones(shp)
            elif len(shp) == 1:
                named[k] = torch.zeros(shp)
            else:
                named[k] = (torch.randn(shp, generator=g) * .02).clamp_(-.04, .04)
        self.load_params(named)

    def step_flops_per_image(self) -> float:
        """Algorithmic FLOPs (2 x MAC) of one training step per image on the LOGICAL pruned shapes (SURVEY 8d formula:
        3 x forward minus the patch-embed data gradient)."""
        N, L, D = self.T, self.L, self.Dv
        f = 2.0 * L * 768 * D + 2.0 * D * self.C
        for H, d, hid in zip(self.heads, self.head_dims, self.hiddens):
            A = H * d
            f += 2.0 * N * D * 3 * A + 2.0 * N * A * D + 4.0 * N * D * hid + 4.0 * H * N * N * d
        return 3.0 * f - 2.0 * L * 768 * D

    def named_parameters(self):
        return {k: self._unpad(k, self.p(k)) for k in self.offsets}

    def named_grads(self):
        return {k: self._unpad(k, self.g(k)) for k in self.offsets}

    def padding_is_clean(self) -> bool:
        """True when every padding entry of the parameter and gradient arenas is exactly zero (the invariant the zero-padded
        layout rests on)."""
        for k in self.offsets:
            for arena in (self.params, self.grads):
                full = self._view(arena, k).clone()
                live = self._pad(k, self._unpad(k, full))
                if not torch.equal(full, live):
                    return False
        return True

    def _fill_hyper(self, lr=None):
        t = self.step_count + 1
        lr = self.lr if lr is None else lr
        h = self._hyper_up.begin()
        for i, (lid, dec) in enumerate(self.groups):
            sc = self.layer_decay ** (self.depth + 1 - lid)
            h[i * 8:i * 8 + 7] = torch.tensor([lr * sc, self.wd if dec else 0.0, 0.9, 0.999, 1e-8, 1 - 0.9 ** t, 1 - 0.999 ** t])

    # ------------------------------------------------------------------------------------------------------------
    def forward(self, images, labels=None, target=None, drop_u=None, train=True):
        """logits + loss. labels int64 [B] (label smoothing) or target fp32 [B, C] (Mixup soft targets).
        train=False: evaluate() (engine.py:222-257) - DropPath off, per-image {cross entropy, top-1, top-5} instead of the loss."""
        B, T, L, M, ML, Dp, Dv = self.B, self.T, self.L, self.M, self.ML, self.Dp, self.Dv
        use_dp = train and self.training_mode and self.drop_path_rate > 0
        if use_dp:
            if drop_u is None:
                drop_u = torch.rand(self.depth * 2, B, device=self.dev)
            ops.droppath_scale(drop_u, self.drop_prob, self.drop_scale)
        ones = self.ones
        ops.patchify(images, self.patches, self.P)
        x0 = self.xs[0]
        ops.gemm(ops.EPI_PATCH, self.patches, self.w("patch_embed.proj.weight"), M=ML, N=Dp, K=768, out0=x0,
                 bias=self.p("patch_embed.proj.bias"), colscale=ones, pos=self.p("pos_embed"),
                 mask_token=self.p("pos_embed"), rowmask=self.zero_mask, tokens=L)
        ops.cls_rows(self.p("cls_token"), self.p("pos_embed"), ones, x0, B, T, Dp)
        for l in range(self.depth):
            pre, a = f"blocks.{l}.", self.blk[l]
            H, A, hid = self.heads[l], self.heads[l] * HD, self.hiddens[l]
            dp1 = self.drop_scale[2 * l] if use_dp else None
            dp2 = self.drop_scale[2 * l + 1] if use_dp else None
            ops.layernorm_fwd(self.xs[l], self.p(pre + "norm1.weight"), self.p(pre + "norm1.bias"), a["x1"], a["mean1"],
                              a["rstd1"], self.eps_ln, d_valid=Dv)
            ops.gemm(ops.EPI_STORE, a["x1"], self.w(pre + "attn.qkv.weight"), M=M, N=3 * A, K=Dp, out0=a["qkv"],
                     bias=self.p(pre + "attn.qkv.bias"))
            ops.attention_fwd(a["qkv"], a["o"], a["lse"], dp1, B, T, H, self.scale)
            # x = x + drop_path(attn(norm1(x))): the residual is the block input, not its normalised copy (vt:158)
            ops.gemm(ops.EPI_STORE, a["o"], self.w(pre + "attn.proj.weight"), M=M, N=Dp, K=A, out0=a["x2"],
                     bias=self.p(pre + "attn.proj.bias"), rowscale=dp1, rows_per_scale=T, bias_rowscaled=True, res=self.xs[l])
            ops.layernorm_fwd(a["x2"], self.p(pre + "norm2.weight"), self.p(pre + "norm2.bias"), a["x3"], a["mean2"],
                              a["rstd2"], self.eps_ln, d_valid=Dv)
            ops.gemm(ops.EPI_FC1, self.w(pre + "mlp.fc1.weight"), a["x3"], M=hid, N=M, K=Dp, out0=a["u"], out1=a["h"],
                     bias=self.p(pre + "mlp.fc1.bias"), colscale=ones, rowscale=dp2, rows_per_scale=T, bn=self.mlp_bn)
            ops.gemm(ops.EPI_STORE, a["h"], self.w(pre + "mlp.fc2.weight"), M=M, N=Dp, K=hid, out0=self.xs[l + 1],
                     bias=self.p(pre + "mlp.fc2.bias"), rowscale=dp2, rows_per_scale=T, bias_rowscaled=True, res=a["x2"],
                     a_mn=True)
        ops.layernorm_fwd(self.xs[self.depth], self.p("norm.weight"), self.p("norm.bias"), self.latent, self.meanf,
                          self.rstdf, self.eps_ln, d_valid=Dv)
        ops.gemm(ops.EPI_STORE, self.latent, self.w("head.weight"), M=B, N=self.C, K=Dp, out0=self.logits, out_fp32=True,
                 bias=self.p("head.bias"), lda=T * Dp)
        if not train:
            ops.eval_metrics(self.logits, labels, self.eval_rows)
            ops.reduce_partials(self.eval_rows, B, 3, self.eval_out, scale=1.0 / B, accumulate=False)
            return self.eval_out
        gs = 1.0 / self.accum_iter
        if target is not None:
            ops.soft_target_cross_entropy(self.logits, target, self.loss_rows, self.dlogits, gs)
        else:
            ops.ls_cross_entropy(self.logits, labels, self.loss_rows, self.dlogits, self.smoothing, gs)
        ops.loss_finalize(self.loss_rows, None, None, None, gs, self.scal)        # scal[0] = mean loss
        return self.scal

    # ------------------------------------------------------------------------------------------------------------
    def backward(self):
        # split-K partials of the weight-gradient GEMMs are finished in fixed order, one reduction launch per block
        with ops.wgrad_batch():
            self._backward()

    def _backward(self):
        B, T, L, M, ML, Dp, Dv = self.B, self.T, self.L, self.M, self.ML, self.Dp, self.Dv
        use_dp = self.training_mode and self.drop_path_rate > 0
        R = self.ln_parts
        ones = self.ones
        # ---- head: only the class-token rows of the final LayerNorm output receive a gradient ----
        dlat = self.gA
        dlat.zero_()
        ops.gemm(ops.EPI_STORE, self.dlogits, self.w("head.weight"), M=B, N=Dp, K=self.C, out0=dlat, ld0=T * Dp, b_mn=True)
        ops.gemm(ops.EPI_WGRAD, self.dlogits, self.latent, M=self.C, N=Dp, K=B, out0=self.g("head.weight"), a_mn=True,
                 b_mn=True, ldb=T * Dp)
        ops.colsum_bf16(self.dlogits, B, self.C, self.g("head.bias"))
        # ---- final LayerNorm: dx is the total gradient of the last block's output -> also its fc2 bias gradient ----
        G = self.gB
        last_dp2 = self.drop_scale[2 * self.depth - 1] if use_dp else None
        ops.layernorm_bwd(dlat, self.xs[self.depth], self.meanf, self.rstdf, self.p("norm.weight"), G, self.pg_, self.pb_,
                          self.pd_, last_dp2, T, d_valid=Dv)
        ops.reduce_partials_multi([(self.pg_, R, Dp, self.g("norm.weight")), (self.pb_, R, Dp, self.g("norm.bias")),
                                   (self.pd_, R, Dp, self.g(f"blocks.{self.depth - 1}.mlp.fc2.bias"))])
        spare = [self.gA, self.gC]
        for l in reversed(range(self.depth)):
            pre, a = f"blocks.{l}.", self.blk[l]
            H, A, hid = self.heads[l], self.heads[l] * HD, self.hiddens[l]
            dp1 = self.drop_scale[2 * l] if use_dp else None
            dp2 = self.drop_scale[2 * l + 1] if use_dp else None
            du, dqkv, dO = self.du[:hid], self.dqkv.view(-1)[:M * 3 * A].view(M, 3 * A), self.dO.view(-1)[:M * A].view(M, A)
            G4 = G                                             # d x_{l+1} (total)
            ops.gemm(ops.EPI_WGRAD, G4, a["h"], M=Dp, N=hid, K=M, out0=self.g(pre + "mlp.fc2.weight"), a_mn=True)
            ops.gemm(ops.EPI_FC2_DGRAD, self.w(pre + "mlp.fc2.weight"), G4, M=hid, N=M, K=Dp, out0=du, aux=a["u"],
                     colscale=ones, rowscale=dp2, rows_per_scale=T, colpart0=self.cp0, colpart1=self.cp1, a_mn=True,
                     bn=self.mlp_bn)
            mlp_jobs = [dict(part=self.cp1, R=self.mlp_parts, N=hid, out=self.g(pre + "mlp.fc1.bias"))]
            ops.gemm(ops.EPI_WGRAD, du, a["x3"], M=hid, N=Dp, K=M, out0=self.g(pre + "mlp.fc1.weight"), b_mn=True)
            G3 = spare.pop()                                   # d LN2 output
            ops.gemm(ops.EPI_STORE, du, self.w(pre + "mlp.fc1.weight"), M=M, N=Dp, K=hid, out0=G3, a_mn=True, b_mn=True)
            G2 = spare.pop()                                   # d x2 (total) = LN2 backward + the residual branch G4
            ops.layernorm_bwd(G3, a["x2"], a["mean2"], a["rstd2"], self.p(pre + "norm2.weight"), G2, self.pg2_, self.pb2_,
                              self.pd2_, dp1, T, dres=G4, d_valid=Dv)
            spare.append(G3)
            spare.append(G4)
            # finished at the end of the block together with the attention / LayerNorm 1 partials (one launch per block)
            mlp_jobs += [(self.pg2_, R, Dp, self.g(pre + "norm2.weight")), (self.pb2_, R, Dp, self.g(pre + "norm2.bias")),
                         (self.pd2_, R, Dp, self.g(pre + "attn.proj.bias"))]
            ops.gemm(ops.EPI_WGRAD, G2, a["o"], M=Dp, N=A, K=M, out0=self.g(pre + "attn.proj.weight"), a_mn=True, b_mn=True)
            ops.gemm(ops.EPI_STORE, G2, self.w(pre + "attn.proj.weight"), M=M, N=A, K=Dp, out0=dO, b_mn=True, rowscale=dp1,
                     rows_per_scale=T)
            ops.attention_bwd(a["qkv"], a["o"], dO, a["lse"], ones, dp1, dqkv, self.att_pg, self.att_pb, B, T, H, self.scale)
            attn_jobs = [dict(part=self.att_pb, R=B, N=3 * A, out=self.g(pre + "attn.qkv.bias"))]
            ops.gemm(ops.EPI_WGRAD, dqkv, a["x1"], M=3 * A, N=Dp, K=M, out0=self.g(pre + "attn.qkv.weight"), a_mn=True,
                     b_mn=True)
            G1 = spare.pop()                                   # d LN1 output
            ops.gemm(ops.EPI_STORE, dqkv, self.w(pre + "attn.qkv.weight"), M=M, N=Dp, K=3 * A, out0=G1, b_mn=True)
            G0 = spare.pop()                                   # d x_l (total) = LN1 backward + the residual branch G2
            has_prev = l > 0
            prev_dp2 = self.drop_scale[2 * l - 1] if (use_dp and has_prev) else None
            ops.layernorm_bwd(G1, self.xs[l], a["mean1"], a["rstd1"], self.p(pre + "norm1.weight"), G0, self.pg_, self.pb_,
                              self.pd_ if has_prev else None, prev_dp2, T, dres=G2, d_valid=Dv)
            spare.append(G1)
            spare.append(G2)
            ln1_jobs = [(self.pg_, R, Dp, self.g(pre + "norm1.weight")), (self.pb_, R, Dp, self.g(pre + "norm1.bias"))]
            if has_prev:
                ln1_jobs.append((self.pd_, R, Dp, self.g(f"blocks.{l - 1}.mlp.fc2.bias")))
            ops.reduce_partials_multi(mlp_jobs + attn_jobs + ln1_jobs)
            ops.wgrad_flush()
            G = G0
        # ---- embedding: x0 = [cls + pos_0 ; conv(patches) + bias + pos_{1..L}] ----
        ops.embed_bwd(G, self.xs[0], ones, self.zero_mask, self.dconv, self.e_gx, self.e_pos, self.e_mt, B, T, Dp)
        ops.reduce_partials_multi([(self.e_pos, 1, T * Dp, self.g("pos_embed")), (self.e_pos, 1, Dp, self.g("cls_token")),
                                   (self.e_pos[1:], L, Dp, self.g("patch_embed.proj.bias"))])
        ops.gemm(ops.EPI_WGRAD, self.dconv, self.patches, M=Dp, N=768, K=ML, out0=self.g("patch_embed.proj.weight"),
                 a_mn=True, b_mn=True)

    # ------------------------------------------------------------------------------------------------------------
    def allreduce_grads(self):
        dp.allreduce_arena(self.grads, self.world, self.pg, self._dp_bounds)

    def optimizer_step(self):
        ops.adamw(self.params, self.grads, self.adam_m, self.adam_v, self.shadow, self.hyper, self._seg_end_c, zero_grad=True)
        self.step_count += 1

    def step(self, images, labels=None, target=None, drop_u=None, update=True, lr=None):
        """One train_one_epoch iteration; returns the device tensor scal (scal[0] = loss)."""
        self._fill_hyper(lr)
        self._hyper_up.upload(self.hyper)
        self.forward(images, labels, target, drop_u)
        self.backward()
        if update:
            self.allreduce_grads()
            self.optimizer_step()
        return self.scal

    def step_graphed(self, images, labels=None, target=None, lr=None):
        """The same step replayed from a CUDA graph (one graph per input-buffer set)."""
        key = (images.data_ptr(), labels.data_ptr() if labels is not None else 0, target.data_ptr() if target is not None else 0)
        self._fill_hyper(lr)
        self._hyper_up.upload(self.hyper)
        entry = self._graphs.get(key)
        if entry is None:
            cur = torch.cuda.current_stream(self.dev)
            side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                self.forward(images, labels, target)
                self.backward()
                self.grads.zero_()
            cur.wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            n0 = ops.LAUNCHES
            with torch.cuda.graph(graph):
                self.forward(images, labels, target)
                self.backward()
                if self.world <= 1:
                    ops.adamw(self.params, self.grads, self.adam_m, self.adam_v, self.shadow, self.hyper, self._seg_end_c,
                              zero_grad=True)
            entry = (graph, ops.LAUNCHES - n0)
            self._graphs[key] = entry
        entry[0].replay()
        ops._count(entry[1])
        if self.world > 1:
            self.allreduce_grads()
            ops.adamw(self.params, self.grads, self.adam_m, self.adam_v, self.shadow, self.hyper, self._seg_end_c,
                      zero_grad=True)
        self.step_count += 1
        return self.scal

    def evaluate(self, images, labels):
        """One evaluate() batch (engine.py:222-257): device tensor [mean cross entropy, top-1 fraction, top-5 fraction]."""
        return self.forward(images, labels, train=False)

    def release_graphs(self):
        self._graphs.clear()
