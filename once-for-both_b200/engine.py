"""SearchStepEngine — one bi-mask DeiT *search training step* (forward, losses, backward, 3x AdamW, optional DP
all-reduce) as an explicit sequence of sm_100a kernels on a flat parameter arena.

Reference step: engine.search_one_epoch loop body (engine.py:95-198) =
  MIMVisionTransformer.forward (vision_transformer.py:614-669, 717-745) -> OFBSearchLOSS (losses.py:80-106) ->
  decoder-loss weighting (engine.py:134-144) -> backward (engine.py:169) -> optimizer_param/arch/decoder.step()
  (optim.py:56-120) [-> DDP all-reduce, search.py:619].

HBM layout
  * parameters, gradients, Adam m / v: four fp32 arenas with identical layout, tensors ordered by optimizer group
    [param no-decay | param decay | decoder no-decay | decoder decay | arch] (search.py:486-559) so one fused AdamW
    launch covers everything; a bf16 shadow arena (same offsets) feeds the GEMMs.
  * activations: bf16, token-major [B*197, C] except the MLP hidden activations u, h, du which are kept transposed
    [hidden, B*197]; per block the backward keeps x_in, LN stats, x1, qkv, lse, o, x2, x3, u, h.
  * nn.Linear weights are used as stored ([out, in]) by forward (K-major B operand), data-gradient (MN-major B operand)
    and weight-gradient GEMMs (MN-major A/B operands) - no transposed copies exist.
There is no CPU path: every op goes through the C ABI in libofb_b200.so.
"""
import ctypes as C
import math
import os
from typing import Dict, List, Optional

import torch

from . import dp, ops
from ._lib import BimaskModule

T_PAD = 8   # arena tensors start at multiples of 8 elements (16 B in bf16, 32 B in fp32)

NO_DECAY_KEYS = ("pos_embed", "cls_token", "dist_token", "scale_weight", "mask_token", "score")
GROUPS = ("param_nd", "param_d", "dec_nd", "dec_d", "arch")


def param_group(name: str, shape) -> str:
    """Optimizer group of a parameter, exactly as search.py:489-508 splits them."""
    if len(shape) == 1 or name.endswith(".bias") or any(k in name for k in NO_DECAY_KEYS):
        return "dec_nd" if "decoder" in name else "param_nd"
    if "alpha" in name:
        return "arch"
    return "dec_d" if "decoder" in name else "param_d"


# ---- search space (layers.py:143-152, 425-462, 813-821) ----
def embed_widths(D):
    return [int((i / D) * D) for i in range(D // 2, D + 1, min(D // 32, 12))]


def head_counts(H):
    return list(range(2, H + 1, 2))


def head_channel_widths(d):
    return [int(d * (i / d)) for i in range(d // 4, d + 1, max(d // 8, 1))]


def hidden_widths(h):
    return [int((i / h) * h) for i in range(h // 4, h + 1, h // 8)]


def flops_polynomial(D, H, d, hid, L, Cn, depth, ae, sd_list, sm_list):
    """(original, searched) FLOPs of MIMVisionTransformer.get_flops (vision_transformer.py:759-783) as a function of
    the weighted-mask sums; works on floats or tensors."""
    n = float(L)
    f_ori = L * D * 768.
    f_s = L * ae * 768.
    for l in range(depth):
        sd, sm = sd_list[l], sm_list[l]
        f_ori += 2 * D * n
        f_s = f_s + 2 * D * n
        f_ori += n * (D * 3 * D) + 3 * n * D + H * n * d * n + H * n * n + 5 * H * n * n + H * n * n * d + n * D * D + n * D
        f_s = f_s + n * (ae * 3 * sd) + 3 * n * sd + n * n * sd + H * n * n + 5 * H * n * n + n * n * sd \
            + n * (sd * ae) + n * ae
        f_ori += (2 * D * hid + D + hid) * n
        f_s = f_s + (ae * sm * 2 + ae + sm) * n
    f_ori += D * Cn
    f_s = f_s + ae * Cn
    return f_ori, f_s


class BimaskTable:
    """Device-side description of every searchable module (order: patch_embed, then attn / mlp per block) plus the
    buffers of the fused bimask_prepare kernels.

    D, H, hidden are the ORIGINAL dims (they stay in the total-FLOPs side of the FLOPs loss). After truncating prune events
    `pruned` carries the current shapes - embed (real width), embed_phys (padded to 8), heads / head_dims / hiddens per block -
    and `spaces[prefix] = (widths, head_counts)` the surviving prefixes of the search-space lists. A pruned attention module
    keeps its heads HD_PHYS = 64 wide physically: score, gate, rank and d gate use that head stride, padding entries stay 0."""

    def __init__(self, D, H, depth, hidden, switches: Dict[str, torch.Tensor], *, w_attn=0.5, w_mlp=0.5, w_embed=0.5,
                 w_flops=5.0, target_flops=1.0, num_classes=1000, num_patches=196, pruned=None, spaces=None):
        self.D, self.H, self.depth, self.hidden = D, H, depth, hidden
        self.d = D // H
        self.L, self.C = num_patches, num_classes
        self.target_flops, self.w_flops = target_flops, w_flops
        self.modules: List[dict] = []
        sw_bytes, widths, gate_off = [], [], 0
        pr = pruned or dict(embed=D, embed_phys=D, heads=[H] * depth, head_dims=[self.d] * depth, hiddens=[hidden] * depth)
        self.D_active = pr["embed"]

        def space(prefix, wj, ni):
            return (list(spaces[prefix][0]), list(spaces[prefix][1])) if spaces is not None else (wj, ni)

        def add(prefix, kind, dim, heads, stride, slot, wj, ni, coef, loss_w):
            nonlocal gate_off
            wj, ni = space(prefix, wj, ni)
            sw = switches[prefix].to(torch.bool).reshape(len(ni) if kind == 2 else 1, len(wj))
            # MAEBlock normalises only the channels whose weighted embed mask is > 0 and concatenates the rest behind them
            # (vision_transformer.py:193-201). Every state compress() can leave keeps the widest surviving embed cell equal to
            # the current width (it truncates first), so all channels are reserved and this engine's all-channel LayerNorm is
            # the same thing; a hand-built switch table whose widest cell is dead would silently diverge - refuse it.
            if kind == 0 and int(sw.sum()) > 1 and not bool(sw[:, -1].any()):
                raise ValueError("patch_embed switch cells: the widest embedding candidate must be alive (the reference would "
                                 "normalise a channel subset, vision_transformer.py:193-201; compress() never leaves this state)")
            self.modules.append(dict(prefix=prefix, kind=kind, dim=dim, heads=heads, stride=stride, slot=slot, n_i=sw.shape[0],
                                     n_j=sw.shape[1], switch_off=len(sw_bytes), width_off=len(widths), gate_off=gate_off,
                                     coef=coef, loss_w=loss_w))
            sw_bytes.extend(int(x) for x in sw.reshape(-1).tolist())
            widths.extend(wj)
            widths.extend(ni)
            gate_off += slot

        add("patch_embed", 0, pr["embed"], 1, pr["embed_phys"], pr["embed_phys"], embed_widths(D), [], 1e-4, w_embed)
        for l in range(depth):
            Hl, dl, hl = pr["heads"][l], pr["head_dims"][l], pr["hiddens"][l]
            add(f"blocks.{l}.attn", 2, dl, Hl, self.d, Hl * self.d, head_channel_widths(self.d), head_counts(H), 4e-4, w_attn)
            add(f"blocks.{l}.mlp", 1, hl, 1, hl, hl, hidden_widths(hidden), [], 1e-4, w_mlp)
        self.total_gate = gate_off
        self._sw_bytes, self._widths = sw_bytes, widths
        self.max_n = max(m["heads"] * m["dim"] for m in self.modules)

    def bind(self, offsets: Dict[str, int], device):
        """offsets: arena offset (in floats) of '<prefix>.alpha' / '<prefix>.score'."""
        n = len(self.modules)
        arr = (BimaskModule * n)()
        for i, m in enumerate(self.modules):
            a = arr[i]
            a.kind, a.dim, a.heads, a.n_i, a.n_j = m["kind"], m["dim"], m["heads"], m["n_i"], m["n_j"]
            a.switch_off, a.width_off, a.gate_off, a.stride = m["switch_off"], m["width_off"], m["gate_off"], m["stride"]
            a.alpha_off, a.score_off = offsets[m["prefix"] + ".alpha"], offsets[m["prefix"] + ".score"]
            a.coef, a.loss_w = m["coef"], m["loss_w"]
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self.mods_dev = raw.to(device)
        self.switches_dev = torch.tensor(self._sw_bytes, dtype=torch.uint8, device=device)
        self.widths_dev = torch.tensor(self._widths, dtype=torch.int32, device=device)
        f32 = dict(dtype=torch.float32, device=device)
        self.gate = torch.zeros(self.total_gate, **f32)
        self.rank = torch.zeros(self.total_gate, dtype=torch.int32, device=device)
        self.aprob = torch.zeros(n * 64, **f32)
        self.wsum = torch.zeros(n, **f32)
        self.sp_loss = torch.zeros(n, **f32)
        self.arch = torch.zeros(8, **f32)
        self.dwsum = torch.zeros(n, **f32)
        return self

    def gate_of(self, i):
        """Physical gate slot of module i (zero at padding entries): what the GEMM / attention epilogues consume."""
        m = self.modules[i]
        return self.gate[m["gate_off"]:m["gate_off"] + m["slot"]]

    def logical(self, i, buf):
        """[heads, dim] view of module i's real entries in a gate-shaped buffer (gate, rank, d gate)."""
        m = self.modules[i]
        return buf[m["gate_off"]:m["gate_off"] + m["heads"] * m["stride"]].view(m["heads"], m["stride"])[:, :m["dim"]] \
            if m["kind"] == 2 else buf[m["gate_off"]:m["gate_off"] + m["dim"]].view(1, m["dim"])

    def pruned_index_sets(self):
        """Kept-unit index sets of every searchable module for every candidate of its search space, i.e. what the reference's
        compress() would physically keep when it slices to that candidate: channels = argsort(score, descending)[:width]
        (per head for attention: layers.py:614-620, 666-670; MLP 932-933, 967-968; embed 268-269, 308-309), heads =
        argsort(sigmoid(score).sum(-1), descending)[:n]. Read from the device-side ranks of the last forward (ties broken by
        the lower index). Returns {prefix: {"channels": {width: [sorted indices per head]}, "heads": {n: sorted indices}}}."""
        rank = self.rank.cpu()
        wl = self._widths
        out = {}
        for i, m in enumerate(self.modules):
            H, dim = m["heads"], m["dim"]
            r = self.logical(i, rank)
            hr, cr = r[:, 0] // dim, r % dim
            widths = wl[m["width_off"]:m["width_off"] + m["n_j"]]
            counts = wl[m["width_off"] + m["n_j"]:m["width_off"] + m["n_j"] + (m["n_i"] if m["kind"] == 2 else 0)]
            out[m["prefix"]] = {
                "channels": {int(w): [torch.nonzero(cr[h] < w).flatten().tolist() for h in range(H)] for w in widths},
                "heads": {int(n): torch.nonzero(hr < n).flatten().tolist() for n in counts},
            }
        return out

    def forward(self, params, w_p_dev):
        n = len(self.modules)
        ops.bimask_fwd(self.mods_dev, n, self.max_n, params, self.switches_dev, self.widths_dev, w_p_dev, self.gate,
                       self.rank, self.aprob, self.wsum, self.sp_loss)
        ops.arch_finalize(self.mods_dev, n, self.wsum, self.sp_loss, self.depth, self.D, self.H, self.d, self.hidden,
                          self.L, self.C, self.target_flops, self.w_flops, self.arch, self.dwsum, d_active=self.D_active)

    def backward(self, params, w_p_dev, dgate, grad_scale, grads):
        ops.bimask_bwd(self.mods_dev, len(self.modules), self.max_n, params, self.switches_dev, self.widths_dev, w_p_dev,
                       dgate, self.rank, self.aprob, self.dwsum, grad_scale, grads)


class SearchStepEngine:
    def __init__(self, embed_dim=384, num_heads=6, depth=12, batch=256, *, mlp_ratio=4, num_classes=1000, img=224,
                 patch=16, drop_path_rate=0.1, lr=1e-3, weight_decay=1e-3, eps_ln=1e-6, smoothing=0.1, w_attn=0.5,
                 w_mlp=0.5, w_embed=0.5, w_flops=5.0, target_flops=1.0, accum_iter=1, warmup_epochs=20, max_ratio=0.95,
                 min_ratio=0.75, device="cuda", switches=None, process_group=None, pruned=None):
        """pruned: shapes of a model that truncating prune events have physically sliced while it is still being searched
        (see from_pruned): dict(embed=D', heads=[H'_l], head_dims=[d'_l], hiddens=[h'_l], spaces={prefix: (widths, head_counts)}).
        embed_dim / num_heads / mlp_ratio stay the ORIGINAL dims (attention scale, total-FLOPs side of the FLOPs loss). The pruned
        tensors live in the zero-padded layout of finetune_engine.py: embedding axis padded to a multiple of 8, heads 64 wide."""
        assert embed_dim % num_heads == 0 and embed_dim // num_heads == 64, "attention kernel is built for head_dim 64"
        self.D0, self.H, self.depth, self.B = embed_dim, num_heads, depth, batch
        self.d = 64
        self.hid = embed_dim * mlp_ratio
        self.Dv = pruned["embed"] if pruned else embed_dim                   # real embedding width
        self.D = (self.Dv + 7) // 8 * 8                                      # physical (padded) embedding width
        self.heads = list(pruned["heads"]) if pruned else [num_heads] * depth
        self.hdims = list(pruned["head_dims"]) if pruned else [64] * depth
        self.hids = list(pruned["hiddens"]) if pruned else [self.hid] * depth
        self.spaces = pruned.get("spaces") if pruned else None
        assert all(0 < x <= 64 and x % 8 == 0 for x in self.hdims) and all(x % 8 == 0 for x in self.hids)
        self.C, self.img, self.P = num_classes, img, patch
        self.L = (img // patch) ** 2
        self.T = self.L + 1
        self.M, self.ML = batch * self.T, batch * self.L
        self.dev = torch.device(device)
        self.lr, self.wd, self.eps_ln, self.smoothing = lr, weight_decay, eps_ln, smoothing
        self.accum_iter = accum_iter
        self.warmup_epochs, self.max_ratio, self.min_ratio = warmup_epochs, max_ratio, min_ratio
        self.scale = float(self.d) ** -0.5
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.step_count = 0
        self.w_p, self.keep_ratio = 0.99, max_ratio
        self._schedule_touched = False
        self.decoder_frozen = False        # post-search phase (enter_post_search)
        self._finish_cache = None
        self.drop_path_rate = drop_path_rate

        # ---- parameter arenas ----
        self.ref_shapes = self._param_shapes()                       # reference (logical) shapes
        shapes = {k: self._padded_shape(k, v) for k, v in self.ref_shapes.items()}      # physical shapes (== logical when unpruned)
        # stable sort: keeps model order inside a group; the bi-mask scores go to the end of their (no-decay) group so that
        # each can be its own AdamW segment (see below)
        order = sorted(shapes, key=lambda k: (GROUPS.index(param_group(k, shapes[k])), k.endswith(".score")))
        self.offsets, self.shapes, off = {}, shapes, 0
        seg_end, cur = [], GROUPS[0]
        for k in order:
            gname = param_group(k, shapes[k])
            if gname != cur:
                while GROUPS[len(seg_end)] != gname:
                    seg_end.append(off)
                cur = gname
            self.offsets[k] = off
            off += (math.prod(shapes[k]) + T_PAD - 1) // T_PAD * T_PAD
        while len(seg_end) < len(GROUPS):
            seg_end.append(off)
        self.n_arena = off
        self.seg_end = seg_end
        # AdamW segments: the four weight / decoder groups, then ONE SEGMENT PER ALPHA TENSOR - a prune event restarts the
        # optimizer state of the alphas it touches, step counter included (optim.AdamW.update(..., initialize=True),
        # optim.py:152-159), so every alpha carries its own bias-correction step
        # ... and so does every score tensor: finalising a module re-initialises its score's state the same way (layers.py:631)
        self.alpha_names = [k for k in order if param_group(k, shapes[k]) == "arch"]
        self.score_names = [k for k in order if k.endswith(".score")]
        tend = lambda k: self.offsets[k] + (math.prod(shapes[k]) + T_PAD - 1) // T_PAD * T_PAD
        self.segments, ends = [("param_nd", None)], [self.offsets[self.score_names[0]]]
        for k in self.score_names:
            self.segments.append(("param_nd", k)); ends.append(tend(k))
        assert ends[-1] == seg_end[0], "scores must close the no-decay group"
        for gi in (1, 2, 3):
            self.segments.append((GROUPS[gi], None)); ends.append(seg_end[gi])
        for k in self.alpha_names:
            self.segments.append(("arch", k)); ends.append(tend(k))
        assert ends[-1] == off and len(ends) <= 64, "AdamW segment table (ofb_b200.h OFB_ADAMW_MAX_SEGMENTS)"
        self._seg_end_c = (C.c_int64 * len(ends))(*ends)
        # optimizer step at which a tensor's Adam state was (re)started by a prune event
        self.alpha_restart = {k: 0 for k in self.alpha_names + self.score_names}
        self._wp_idx = len(self.segments) * 8
        f32 = dict(dtype=torch.float32, device=self.dev)
        bf = dict(dtype=torch.bfloat16, device=self.dev)
        self.params = torch.zeros(off, **f32)
        self.grads = torch.zeros(off, **f32)
        self.adam_m = torch.zeros(off, **f32)
        self.adam_v = torch.zeros(off, **f32)
        self.shadow = torch.zeros(off, **bf)
        self.alpha_patch = torch.ones(1, 1, **f32)          # frozen when --patch_search is off
        nh = self._wp_idx + 8
        self._hyper_up = ops.HyperUploader(nh, self.dev)
        self.hyper = torch.zeros(nh, **f32)                    # [segments x 8 ..., w_p]

        # ---- bi-mask ----
        if switches is None:
            assert pruned is None, "a pruned engine needs the surviving switch cells"
            switches = {"patch_embed": torch.ones(1, len(embed_widths(self.D0)), dtype=torch.bool)}
            for l in range(depth):
                switches[f"blocks.{l}.attn"] = torch.ones(len(head_counts(self.H)), len(head_channel_widths(self.d)),
                                                          dtype=torch.bool)
                switches[f"blocks.{l}.mlp"] = torch.ones(1, len(hidden_widths(self.hid)), dtype=torch.bool)
        self.switches = switches
        # finished embedding search -> standard pre-norm blocks (vision_transformer.py:193, 203-204); otherwise the residual is
        # taken from the normalised input (the search-mode quirk)
        self.prenorm = int(switches["patch_embed"].sum()) == 1
        self._loss_w = dict(w_attn=w_attn, w_mlp=w_mlp, w_embed=w_embed, w_flops=w_flops, target_flops=target_flops)
        self.bimask = BimaskTable(self.D0, self.H, depth, self.hid, switches, w_attn=w_attn, w_mlp=w_mlp, w_embed=w_embed,
                                  w_flops=w_flops, target_flops=target_flops, num_classes=num_classes, num_patches=self.L,
                                  pruned=dict(embed=self.Dv, embed_phys=self.D, heads=self.heads, head_dims=self.hdims,
                                              hiddens=self.hids),
                                  spaces=self.spaces).bind(self.offsets, self.dev)
        self.dgate = torch.zeros(self.bimask.total_gate, **f32)

        # ---- activations ----
        M, ML, D, hid, B, T, H = self.M, self.ML, self.D, self.hid, self.B, self.T, self.H
        self.patches = torch.empty(ML, 3 * patch * patch, **bf)
        self.mask = torch.zeros(B, self.L, **f32)
        self.tgt = torch.zeros(ML, 768, **f32)
        dpr = torch.linspace(0, drop_path_rate, depth).repeat_interleave(2)
        self.drop_prob = dpr.to(self.dev)
        self.drop_scale = torch.ones(depth * 2, B, **f32)
        # MLP hidden activations are kept TRANSPOSED, [hidden, tokens] with the token pitch padded to 16 bytes: the fc1 /
        # fc2-dgrad epilogue threads then own one hidden unit each (gate, bias and their gradients are per-thread scalars)
        self.ldT = (M + 7) // 8 * 8
        self.mlp_bn = int(os.environ.get("OFB_MLP_BN", "256"))      # token tile of the transposed-hidden GEMMs
        self.bn_nD = 0          # N tile of the N = embed_dim GEMMs (0 = the library's cost model); tools/step_breakdown.py A/Bs it
        self.mlp_parts = ops.gemm_mlp_partial_rows(M, self.mlp_bn)
        self.xs = [torch.zeros(M, D, **bf) for _ in range(depth + 1)]      # xs[l] = input of block l; xs[depth] = output
        self.blk = []
        for l in range(depth):
            A, hl = self.heads[l] * 64, self.hids[l]
            self.blk.append(dict(
                mean1=torch.empty(M, **f32), rstd1=torch.empty(M, **f32), x1=torch.empty(M, D, **bf),
                qkv=torch.empty(M, 3 * A, **bf), lse=torch.empty(B, self.heads[l], T, **f32), o=torch.empty(M, A, **bf),
                x2=torch.empty(M, D, **bf), mean2=torch.empty(M, **f32), rstd2=torch.empty(M, **f32),
                x3=torch.empty(M, D, **bf), u=torch.empty(hl, self.ldT, **bf), h=torch.empty(hl, self.ldT, **bf)))
        Amax, hid = max(self.heads) * 64, max(self.hids)
        self.meanf, self.rstdf = torch.empty(M, **f32), torch.empty(M, **f32)
        self.latent = torch.empty(M, D, **bf)
        self.logits = torch.empty(B, self.C, **f32)
        self.loss_rows = torch.empty(B, **f32)
        self.dlogits = torch.empty(B, self.C, **bf)
        self.sgn = torch.zeros(M, 768, **bf)
        self.dec_bn = 256
        self.dec_part = torch.zeros(((M + 127) // 128) * (768 // self.dec_bn) * 8, **f32)   # one partial per epilogue warp
        self.scal = torch.zeros(8, **f32)
        self.zero_mask = torch.zeros(B, self.L, **f32)           # eval mode: no PMIM masking
        self.eval_rows, self.eval_out = torch.zeros(B, 3, **f32), torch.zeros(3, **f32)
        # backward scratch
        self.gA, self.gB, self.gC = (torch.empty(M, D, **bf) for _ in range(3))
        self.du = torch.empty(hid, self.ldT, **bf)
        self.dqkv = torch.empty(M, 3 * Amax, **bf)
        self.dObuf = torch.empty(M, Amax, **bf)
        self.dconv = torch.empty(ML, D, **bf)
        self.ln_parts = ops.layernorm_bwd_parts(M)
        self.pg_, self.pb_, self.pd_ = (torch.empty(self.ln_parts, D, **f32) for _ in range(3))
        # second set: the partials of a block's LayerNorm 2 are finished together with those of its LayerNorm 1 (one launch per block)
        self.pg2_, self.pb2_, self.pd2_ = (torch.empty(self.ln_parts, D, **f32) for _ in range(3))
        mt = (M + 127) // 128
        self.cp0, self.cp1 = torch.empty(self.mlp_parts, hid, **f32), torch.empty(self.mlp_parts, hid, **f32)
        self.att_pg, self.att_pb = torch.empty(B, Amax, **f32), torch.empty(B, 3 * Amax, **f32)
        self.e_gx, self.e_pos, self.e_mt = (torch.empty(T, D, **f32) for _ in range(3))
        self.rand_u = torch.empty(B * self.L + depth * 2 * B, **f32)
        # the exchange that follows backward is ONE all-reduce of the arena: on 8 x B200 89.6 MB take 0.30 ms in one piece, 0.46 ms
        # in four (profiles/r02b_allreduce_8gpu.txt); OFB_DP_BUCKETS restores a bucketed exchange
        self._dp_bounds = dp.bucket_bounds(self.n_arena, max_buckets=int(os.environ.get("OFB_DP_BUCKETS", "1")))
        # exchange overlapped with backward (OFB_DP_OVERLAP=1): buckets of whole blocks' weight gradients (contiguous in the
        # decay group) are all-reduced as soon as backward has passed them. Default is the single exchange after backward:
        # measured on 2 x B200 (profiles/r01c_dp_overlap_ab.txt) the overlapped form is 0.1-0.2 ms/step SLOWER - every
        # compute kernel is persistent with one CTA per SM and 200+ KB of shared memory, so the NCCL CTAs cannot co-reside
        # and instead delay the CTAs of the next compute kernel on the SMs they hold - while the whole exchange costs 0.2 ms
        self.dp_overlap = self.world > 1 and os.environ.get("OFB_DP_OVERLAP", "0") == "1"
        blk_rng = []
        for l in range(depth):
            lo = self.offsets[f"blocks.{l}.attn.qkv.weight"]
            k = f"blocks.{l}.mlp.fc2.weight"
            hi = self.offsets[k] + (math.prod(shapes[k]) + T_PAD - 1) // T_PAD * T_PAD
            assert not blk_rng or blk_rng[-1][1] == lo, "block weight runs must be adjacent in the arena"
            blk_rng.append((lo, hi))
        early, tail = dp.overlap_plan(self.n_arena, blk_rng, int(os.environ.get("OFB_DP_BLOCKS_PER_BUCKET", "2")),
                                      int(os.environ.get("OFB_DP_TAIL_BLOCKS", "2")))
        self._reducer = dp.OverlappedReducer(self.grads, self.world, self.pg, early, tail)
        # the single exchange after backward as nodes of the step graph (no host launches between backward and AdamW)
        self.dp_in_graph = self.world > 1 and os.environ.get("OFB_DP_GRAPH", "0") == "1"
        self._graphs = {}          # (images ptr, labels ptr, keep, target ptr, update, pinned noise) -> (CUDAGraph, launches per replay)
        self._noise_in = None      # persistent PMIM-noise input of graphs captured with a pinned draw
        self._side = None          # side stream of the gate construction (see forward)
        self._side2 = None         # side stream of the PMIM target normalisation

    # ------------------------------------------------------------------------------------------------------------
    def _space(self, prefix, widths, counts):
        return (self.spaces[prefix][0], self.spaces[prefix][1]) if self.spaces is not None else (widths, counts)

    def _param_shapes(self):
        """Reference state_dict names / shapes of the (possibly truncated) search model, in named_parameters() order."""
        D, L, Cn, H0, d0 = self.Dv, self.L, self.C, self.H, self.d
        s = {"cls_token": (1, 1, D), "pos_embed": (1, L + 1, D), "mask_token": (1, 1, D),
             "patch_embed.alpha": (1, len(self._space("patch_embed", embed_widths(self.D0), [])[0])), "patch_embed.score": (1, D),
             "patch_embed.proj.weight": (D, 3, self.P, self.P), "patch_embed.proj.bias": (D,)}
        for l in range(self.depth):
            p = f"blocks.{l}."
            H, d, hid = self.heads[l], self.hdims[l], self.hids[l]
            wj, ni = self._space(p + "attn", head_channel_widths(d0), head_counts(H0))
            s[p + "norm1.weight"] = (D,); s[p + "norm1.bias"] = (D,)
            s[p + "attn.alpha"] = (len(ni), len(wj))
            s[p + "attn.score"] = (H, d)
            s[p + "attn.qkv.weight"] = (3 * H * d, D); s[p + "attn.qkv.bias"] = (3 * H * d,)
            s[p + "attn.proj.weight"] = (D, H * d); s[p + "attn.proj.bias"] = (D,)
            s[p + "norm2.weight"] = (D,); s[p + "norm2.bias"] = (D,)
            s[p + "mlp.alpha"] = (1, len(self._space(p + "mlp", hidden_widths(self.hid), [])[0])); s[p + "mlp.score"] = (1, hid)
            s[p + "mlp.fc1.weight"] = (hid, D); s[p + "mlp.fc1.bias"] = (hid,)
            s[p + "mlp.fc2.weight"] = (D, hid); s[p + "mlp.fc2.bias"] = (D,)
        s["norm.weight"] = (D,); s["norm.bias"] = (D,)
        s["head.weight"] = (Cn, D); s["head.bias"] = (Cn,)
        s["decoder.0.weight"] = (768, D, 1, 1); s["decoder.0.bias"] = (768,)
        return s

    def _blk(self, name):
        return int(name.split(".")[1])

    def _padded_shape(self, name, shp):
        """Physical (arena) shape of a tensor: embedding axis padded to self.D, attention heads 64 wide."""
        Dv, Dp = self.Dv, self.D
        if name.endswith(".alpha") or name.endswith("mlp.score") or name.endswith("mlp.fc1.bias") or name in ("head.bias",
                                                                                                           "decoder.0.bias"):
            return tuple(shp)
        if name.endswith("attn.score"):
            return (shp[0], 64)
        if name.endswith("attn.qkv.weight"):
            return (3 * self.heads[self._blk(name)] * 64, Dp)
        if name.endswith("attn.qkv.bias"):
            return (3 * self.heads[self._blk(name)] * 64,)
        if name.endswith("attn.proj.weight"):
            return (Dp, self.heads[self._blk(name)] * 64)
        return tuple(Dp if (x == Dv and i == self._embed_axis(name, len(shp))) else x for i, x in enumerate(shp))

    @staticmethod
    def _embed_axis(name, rank):
        """Axis of the embedding dimension in a tensor that is padded along it only."""
        if name in ("cls_token", "pos_embed", "mask_token", "patch_embed.score", "head.weight") or name.endswith("mlp.fc1.weight"):
            return rank - 1
        if name == "decoder.0.weight":
            return 1
        return 0          # vectors over the embedding axis, patch_embed.proj.weight, fc2.weight

    def _pad(self, name, t):
        """reference-shaped tensor -> zero-padded physical layout."""
        phys = self.shapes[name]
        if tuple(t.shape) == tuple(phys):
            return t
        out = torch.zeros(phys, dtype=t.dtype, device=t.device)
        if name.endswith("attn.qkv.weight") or name.endswith("attn.qkv.bias") or name.endswith("attn.proj.weight"):
            l = self._blk(name)
            H, d, Dv, Dp = self.heads[l], self.hdims[l], self.Dv, self.D
            if name.endswith("qkv.weight"):
                out.view(3, H, 64, Dp)[:, :, :d, :Dv] = t.reshape(3, H, d, Dv)
            elif name.endswith("qkv.bias"):
                out.view(3, H, 64)[:, :, :d] = t.reshape(3, H, d)
            else:
                out.view(Dp, H, 64)[:Dv, :, :d] = t.reshape(Dv, H, d)
            return out
        out[tuple(slice(0, x) for x in t.shape)] = t
        return out

    def _unpad(self, name, t):
        """physical layout -> reference-shaped tensor."""
        ref = self.ref_shapes[name]
        if tuple(t.shape) == tuple(ref):
            return t
        if name.endswith("attn.qkv.weight") or name.endswith("attn.qkv.bias") or name.endswith("attn.proj.weight"):
            l = self._blk(name)
            H, d, Dv, Dp = self.heads[l], self.hdims[l], self.Dv, self.D
            if name.endswith("qkv.weight"):
                return t.view(3, H, 64, Dp)[:, :, :d, :Dv].reshape(ref)
            if name.endswith("qkv.bias"):
                return t.view(3, H, 64)[:, :, :d].reshape(ref)
            return t.view(Dp, H, 64)[:Dv, :, :d].reshape(ref)
        return t[tuple(slice(0, x) for x in ref)]

    def _view(self, arena, name):
        o, shp = self.offsets[name], self.shapes[name]
        return arena[o:o + math.prod(shp)].view(shp)

    def p(self, name):
        return self._view(self.params, name)

    def g(self, name):
        return self._view(self.grads, name)

    def w(self, name):
        """bf16 shadow weight as a 2-D [out, in] matrix."""
        v = self._view(self.shadow, name)
        return v.reshape(v.shape[0], -1)

    def named_parameters(self):
        """Parameters in reference shapes (views of the arena when the model is unpruned)."""
        return {k: self._unpad(k, self.p(k)) for k in self.offsets}

    def named_grads(self):
        return {k: self._unpad(k, self.g(k)) for k in self.offsets}

    def load_params(self, named: Dict[str, torch.Tensor]):
        for k in self.offsets:
            self.p(k).copy_(self._pad(k, named[k].to(self.dev, torch.float32).reshape(self.ref_shapes[k])))
        self.sync_shadow()

    def padding_is_clean(self) -> bool:
        """Every padding entry of the parameter and gradient arenas is exactly zero (invariant of the pruned layout)."""
        for k in self.offsets:
            for arena in (self.params, self.grads):
                full = self._view(arena, k)
                if not torch.equal(full, self._pad(k, self._unpad(k, full).clone())):
                    return False
        return True

    def init_params(self, seed=0):
        """Reference-style initialisation (trunc-normal .02 weights, zero biases, alpha~U(0,1), score~trunc-normal .2:
        vision_transformer.py:497-519, layers.py:147-155, 455-467, 817-824)."""
        g = torch.Generator(device="cpu").manual_seed(seed)
        for k, shp in self.ref_shapes.items():
            t = torch.zeros(shp)
            if k.endswith("alpha"):
                t = torch.rand(shp, generator=g)
            elif k.endswith("score"):
                t = (torch.randn(shp, generator=g) * .2).clamp_(-2, 2)
            elif k.endswith("norm1.weight") or k.endswith("norm2.weight") or k == "norm.weight":
                t = torch.ones(shp)
            elif k.endswith(".bias"):
                t = torch.zeros(shp)
            elif k == "patch_embed.proj.weight":
                bound = math.sqrt(6.0 / (768 + self.Dv))
                t = (torch.rand(shp, generator=g) * 2 - 1) * bound
            else:
                t = (torch.randn(shp, generator=g) * .02).clamp_(-2, 2)
            self.p(k).copy_(self._pad(k, t))
        self.sync_shadow()

    def pruned_index_sets(self):
        """Per-layer kept-unit index sets after thresholding (see BimaskTable.pruned_index_sets); valid after a forward."""
        return self.bimask.pruned_index_sets()

    def sync_shadow(self):
        ops.cast_bf16(self.params, self.shadow)

    # ------------------------------------------------------------------------------------------------------------
    def set_schedule(self, epoch_frac: float):
        """update_w (layers.py:484-486) and adjust_masking_ratio (vision_transformer.py:521-523)."""
        if epoch_frac > self.warmup_epochs:
            # both hooks only act while `epoch <= warmup_epochs`; later the last value stands - which is also what keeps
            # reset_mask_ratio(1.0) of the post-search phase in force
            if not self._schedule_touched:
                epoch_frac = float(self.warmup_epochs)        # a run that starts past the warm-up: the warm-up end values
            else:
                return
        self._schedule_touched = True
        e = epoch_frac
        self.w_p = (0.1 - 0.99) / self.warmup_epochs * e + 0.99
        self.keep_ratio = self.max_ratio - (self.max_ratio - self.min_ratio) * e / self.warmup_epochs

    @property
    def finish_search(self) -> bool:
        """Every searchable module is down to one cell: compress() returned finish_search=True (vision_transformer.py:785-950).
        From then on the criterion returns the base loss alone (losses.py:105-106) and there is no architecture optimizer
        (engine.py:206-208): the architecture loss is left out of the total and the alpha segments are not updated."""
        if self._finish_cache is None:
            self._finish_cache = all(int(sw.sum()) == 1 for sw in self.switches.values())
        return self._finish_cache

    def enter_post_search(self):
        """The epoch-boundary switch of search.py:641-656 once the search has finished: reset_mask_ratio(1.0) (no PMIM masking
        -> no decoder branch, vision_transformer.py:595-612, 719), freeze_decoder() (mask_token and the decoder get no gradient
        and no update, vt:534-539; optimizer_decoder = None). The caller feeds Mixup soft targets from then on
        (step(..., target=...), SoftTargetCrossEntropy, search.py:651-655)."""
        assert self.finish_search, "search.py:642 enters this phase only with finish_search"
        self.keep_ratio = 1.0
        self._schedule_touched = True
        self.decoder_frozen = True
        # a frozen parameter is skipped by the optimizer (grad None, optim.py:70-71). Its gradient stays exactly zero here; with
        # zeroed moments AdamW then leaves it bit-identical (0 / (0 + eps)), the decoder groups additionally get lr 0
        for k in ("mask_token", "decoder.0.weight", "decoder.0.bias"):
            self._view(self.adam_m, k).zero_()
            self._view(self.adam_v, k).zero_()
        self.release_graphs()

    def _fill_hyper(self, lrs=None):
        lrs = lrs or {}
        h = self._hyper_up.begin()
        fin = self.finish_search
        for i, (gname, alpha_name) in enumerate(self.segments):
            t = self.step_count + 1 - (self.alpha_restart[alpha_name] if alpha_name else 0)
            b1, b2 = (0.5, 0.999) if gname == "arch" else (0.9, 0.999)
            lr = lrs.get(gname, self.lr)
            wd = 0.0 if gname.endswith("_nd") else self.wd
            # optimizer_arch = None after finish_search (engine.py:206-208); optimizer_decoder = None in the post-search phase;
            # the alpha of a module that is already down to one cell left optimizer_arch when it was finalised (requires_grad
            # False and removed from the param group, optim.py:179-182): no decay either
            done = gname == "arch" and int(self.switches[alpha_name[:-len(".alpha")]].sum()) == 1
            if (gname == "arch" and fin) or done or (gname.startswith("dec_") and self.decoder_frozen):
                lr = wd = 0.0
            h[i * 8:i * 8 + 7] = torch.tensor([lr, wd, b1, b2, 1e-8, 1 - b1 ** t, 1 - b2 ** t])
        h[self._wp_idx] = self.w_p

    # ------------------------------------------------------------------------------------------------------------
    def forward(self, images, labels, noise=None, drop_u=None, train=True, target=None, mix=None):
        """Forward + losses. images fp32 [B,3,224,224] (device), labels int64 [B].
        target: fp32 [B, C] soft targets (Mixup, post-search phase) -> timm SoftTargetCrossEntropy instead of label smoothing.
        mix: mixup.MixParams - `images` is the UNMIXED batch and the Mixup / CutMix blend is fused into the im2col (only when PMIM
        is off: nothing else reads the images then).
        With keep_ratio == 1 (enter_post_search) PMIM is off: no masking, no target normalisation, no decoder GEMM, decoder
        loss 0 (vision_transformer.py:595-612, 719-730).
        train=False is the reference's eval mode while the search is running (engine.evaluate, engine.py:222-257 ->
        MIMVisionTransformer.forward with self.training False): same gates, no PMIM masking (vt:631-638), DropPath identity,
        no decoder branch (vt:719), and instead of the training losses the per-image {cross entropy, top-1, top-5} rows."""
        B, D, H, T, L, M, ML, hid = self.B, self.D, self.H, self.T, self.L, self.M, self.ML, self.hid
        bm = self.bimask
        w_p_dev = self.hyper[self._wp_idx:self._wp_idx + 1]
        # the gate construction (two small latency-bound kernels) runs on a side stream next to the image-side preparation
        # (PMIM mask, patchify, target normalisation); it joins before the patch-embed GEMM, the first consumer of a gate.
        # Under CUDA-graph capture this becomes a parallel branch of the graph.
        cur = torch.cuda.current_stream(self.dev)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.dev)
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            bm.forward(self.params, w_p_dev)
        # random draws stay in PyTorch (RNG parity with the reference's torch.rand), masks are built by our kernels
        keep = int(L * self.keep_ratio)
        pmim = train and keep != L                    # `len_keeps != [L]`, vision_transformer.py:595
        self._pmim = pmim
        assert mix is None or mix.identity or not pmim, "fused Mixup needs PMIM off; mix the batch with mixup.Mixup first"
        if train:
            if pmim:
                if noise is None:
                    noise = torch.rand(B, L, device=self.dev)
                ops.pmim_mask(noise, self.mask, keep)
            if drop_u is None:
                drop_u = torch.rand(self.depth * 2, B, device=self.dev)
            ops.droppath_scale(drop_u, self.drop_prob, self.drop_scale)
        rowmask = self.mask if pmim else self.zero_mask
        if pmim:
            # the PMIM targets (local normalisation of the masked patches) are only consumed by the decoder GEMM at the very
            # end of forward: a second branch, filling the tails of the block kernels instead of sitting on the critical path
            if self._side2 is None:
                self._side2 = torch.cuda.Stream(device=self.dev)
            side_targets = os.environ.get("OFB_SIDE_TARGETS", "1") == "1"
            if side_targets:
                self._side2.wait_stream(cur)
                with torch.cuda.stream(self._side2):
                    ops.norm_targets(images, self.mask, self.tgt)
            else:
                ops.norm_targets(images, self.mask, self.tgt)
        if mix is not None and not mix.identity:
            ops.patchify_mixup(images, self.patches, mix.lam, mix.box, self.P)
        else:
            ops.patchify(images, self.patches, self.P)
        cur.wait_stream(self._side)
        g_e = bm.gate_of(0)
        x0 = self.xs[0]
        ops.gemm(ops.EPI_PATCH, self.patches, self.w("patch_embed.proj.weight"), M=ML, N=D, K=768, out0=x0,
                 bias=self.p("patch_embed.proj.bias"), colscale=g_e, pos=self.p("pos_embed"),
                 mask_token=self.p("mask_token"), rowmask=rowmask, tokens=L)
        ops.cls_rows(self.p("cls_token"), self.p("pos_embed"), g_e, x0, B, T, D)
        Dv = self.Dv
        for l in range(self.depth):
            pre, a = f"blocks.{l}.", self.blk[l]
            H, A, hid = self.heads[l], self.heads[l] * 64, self.hids[l]     # this block's heads, physical qkv width, hidden width
            g_a, g_m = bm.gate_of(1 + 2 * l), bm.gate_of(2 + 2 * l)
            dp1, dp2 = (self.drop_scale[2 * l], self.drop_scale[2 * l + 1]) if train else (None, None)
            ops.layernorm_fwd(self.xs[l], self.p(pre + "norm1.weight"), self.p(pre + "norm1.bias"), a["x1"], a["mean1"],
                              a["rstd1"], self.eps_ln, d_valid=Dv)
            ops.gemm(ops.EPI_STORE, a["x1"], self.w(pre + "attn.qkv.weight"), M=M, N=3 * A, K=D, out0=a["qkv"],
                     bias=self.p(pre + "attn.qkv.bias"), colscale=g_a, colscale_period=A)
            ops.attention_fwd(a["qkv"], a["o"], a["lse"], dp1, B, T, H, self.scale)
            ops.gemm(ops.EPI_STORE, a["o"], self.w(pre + "attn.proj.weight"), M=M, N=D, K=A, out0=a["x2"],
                     bias=self.p(pre + "attn.proj.bias"), rowscale=dp1, rows_per_scale=T, bias_rowscaled=True,
                     res=self.xs[l] if self.prenorm else a["x1"], bn=self.bn_nD)
            ops.layernorm_fwd(a["x2"], self.p(pre + "norm2.weight"), self.p(pre + "norm2.bias"), a["x3"], a["mean2"],
                              a["rstd2"], self.eps_ln, d_valid=Dv)
            # fc1 computes the transposed hidden activations u^T, h^T = [hidden, tokens] (weight is the M operand)
            ops.gemm(ops.EPI_FC1, self.w(pre + "mlp.fc1.weight"), a["x3"], M=hid, N=M, K=D, out0=a["u"], out1=a["h"],
                     bias=self.p(pre + "mlp.fc1.bias"), colscale=g_m, rowscale=dp2, rows_per_scale=T, bn=self.mlp_bn)
            ops.gemm(ops.EPI_STORE, a["h"], self.w(pre + "mlp.fc2.weight"), M=M, N=D, K=hid, out0=self.xs[l + 1],
                     bias=self.p(pre + "mlp.fc2.bias"), rowscale=dp2, rows_per_scale=T, bias_rowscaled=True,
                     res=a["x2"] if self.prenorm else a["x3"], a_mn=True, bn=self.bn_nD)
        ops.layernorm_fwd(self.xs[self.depth], self.p("norm.weight"), self.p("norm.bias"), self.latent, self.meanf,
                          self.rstdf, self.eps_ln, d_valid=Dv)
        # head on the cls rows (row stride T*D), label-smoothing CE
        ops.gemm(ops.EPI_STORE, self.latent, self.w("head.weight"), M=B, N=self.C, K=D, out0=self.logits, out_fp32=True,
                 bias=self.p("head.bias"), lda=T * D)
        if not train:
            ops.eval_metrics(self.logits, labels, self.eval_rows)
            ops.reduce_partials(self.eval_rows, B, 3, self.eval_out, scale=1.0 / B, accumulate=False)
            return self.eval_out
        gs = 1.0 / self.accum_iter
        if target is not None:
            ops.soft_target_cross_entropy(self.logits, target, self.loss_rows, self.dlogits, gs)
        else:
            ops.ls_cross_entropy(self.logits, labels, self.loss_rows, self.dlogits, self.smoothing, gs)
        arch = None if self.finish_search else bm.arch       # finish_search: the criterion returns the base loss alone
        if not pmim:
            ops.loss_finalize(self.loss_rows, None, None, arch, gs, self.scal)
            return self.scal
        # PMIM decoder + masked L1 against the locally normalised pixels
        if side_targets:
            cur.wait_stream(self._side2)
        ops.gemm(ops.EPI_DECODER, self.latent, self.w("decoder.0.weight"), M=M, N=768, K=D, out0=self.sgn,
                 bias=self.p("decoder.0.bias"), rowmask=self.mask, target=self.tgt, tokens=L, colpart0=self.dec_part,
                 bn=self.dec_bn)
        ops.loss_finalize(self.loss_rows, self.dec_part, self.mask, arch, gs, self.scal)
        return self.scal

    # ------------------------------------------------------------------------------------------------------------
    def backward(self, exchange=False, loss_grads=True):
        """Backward of the step. exchange=True: the data-parallel bucket all-reduces are launched from inside (each bucket
        as soon as backward has passed its blocks) and joined at the end, so the gradients are averaged on return.
        loss_grads=False: d score / d alpha carry the network path only - the sparsity / FLOPs loss terms are left to the caller
        (modules.py: under the reference's own training loop OFBSearchLOSS computes them with autograd)."""
        # the split-K partials of the weight-gradient GEMMs are finished in fixed order, one reduction launch per block
        with ops.wgrad_batch():
            self._backward(exchange, loss_grads)

    def _backward(self, exchange, loss_grads):
        B, D, H, T, L, M, ML, hid = self.B, self.D, self.H, self.T, self.L, self.M, self.ML, self.hid
        bm = self.bimask
        red = self._reducer if (exchange and self.world > 1) else None
        if red is not None:
            red.begin()
        gs = 1.0 / self.accum_iter
        dec_scale = self.scal[5:6]
        R = self.ln_parts
        mt = (M + 127) // 128

        # ---- decoder + head ----
        dlat = self.gA
        pmim = self._pmim
        if pmim:
            ops.gemm(ops.EPI_STORE, self.sgn, self.w("decoder.0.weight"), M=M, N=D, K=768, out0=dlat, b_mn=True,
                     scale_ptr=dec_scale)
        else:
            dlat.zero_()                   # PMIM off: only the cls rows of the latent carry a gradient (the head's)
        ops.gemm(ops.EPI_STORE, self.dlogits, self.w("head.weight"), M=B, N=D, K=self.C, out0=dlat, ld0=T * D, b_mn=True)
        if pmim and not self.decoder_frozen:
            ops.gemm(ops.EPI_WGRAD, self.sgn, self.latent, M=768, N=D, K=M, out0=self.g("decoder.0.weight").view(768, D),
                     a_mn=True, b_mn=True, scale_ptr=dec_scale)
            ops.colsum_bf16(self.sgn, M, 768, self.g("decoder.0.bias"), scale_dev=dec_scale)
        ops.gemm(ops.EPI_WGRAD, self.dlogits, self.latent, M=self.C, N=D, K=B, out0=self.g("head.weight"), a_mn=True,
                 b_mn=True, ldb=T * D)
        ops.colsum_bf16(self.dlogits, B, self.C, self.g("head.bias"))

        # ---- final LayerNorm ----
        last_dp2 = self.drop_scale[2 * self.depth - 1]
        G = self.gB
        Dv = self.Dv
        ops.layernorm_bwd(dlat, self.xs[self.depth], self.meanf, self.rstdf, self.p("norm.weight"), G, self.pg_,
                          self.pb_, self.pd_, last_dp2, T, d_valid=Dv)
        ops.reduce_partials_multi([(self.pg_, R, D, self.g("norm.weight")), (self.pb_, R, D, self.g("norm.bias")),
                                   (self.pd_, R, D, self.g(f"blocks.{self.depth - 1}.mlp.fc2.bias"))])
        spare = [self.gA, self.gC]

        for l in reversed(range(self.depth)):
            pre, a = f"blocks.{l}.", self.blk[l]
            H, A, hid = self.heads[l], self.heads[l] * 64, self.hids[l]
            du = self.du[:hid]
            dqkv = self.dqkv.view(-1)[:M * 3 * A].view(M, 3 * A)
            dO = self.dObuf.view(-1)[:M * A].view(M, A)
            i_a, i_m = 1 + 2 * l, 2 + 2 * l
            g_a, g_m = bm.gate_of(i_a), bm.gate_of(i_m)
            dp1, dp2 = self.drop_scale[2 * l], self.drop_scale[2 * l + 1]
            G4 = G
            # fc2: weight grad (h already carries DropPath), data grad fused with GELU' / gate / column partials
            ops.gemm(ops.EPI_WGRAD, G4, a["h"], M=D, N=hid, K=M, out0=self.g(pre + "mlp.fc2.weight"), a_mn=True)
            ops.gemm(ops.EPI_FC2_DGRAD, self.w(pre + "mlp.fc2.weight"), G4, M=hid, N=M, K=D, out0=du, aux=a["u"],
                     colscale=g_m, rowscale=dp2, rows_per_scale=T, colpart0=self.cp0, colpart1=self.cp1, a_mn=True,
                     bn=self.mlp_bn)
            m_off = bm.modules[i_m]["gate_off"]
            mlp_jobs = [dict(part=self.cp0, R=self.mlp_parts, N=hid, out=self.dgate[m_off:m_off + hid], accumulate=False),
                        dict(part=self.cp1, R=self.mlp_parts, N=hid, out=self.g(pre + "mlp.fc1.bias"))]
            ops.gemm(ops.EPI_WGRAD, du, a["x3"], M=hid, N=D, K=M, out0=self.g(pre + "mlp.fc1.weight"), b_mn=True)
            G3 = spare.pop()
            pn = self.prenorm         # pre-norm: the residual gradient bypasses the LayerNorm and joins inside its backward
            ops.gemm(ops.EPI_STORE, du, self.w(pre + "mlp.fc1.weight"), M=M, N=D, K=hid, out0=G3, a_mn=True, b_mn=True,
                     res=None if pn else G4, bn=self.bn_nD)
            if not pn:
                spare.append(G4)
            # LayerNorm 2 (+ proj bias grad)
            G2 = spare.pop()
            ops.layernorm_bwd(G3, a["x2"], a["mean2"], a["rstd2"], self.p(pre + "norm2.weight"), G2, self.pg2_, self.pb2_,
                              self.pd2_, dp1, T, d_valid=Dv, dres=G4 if pn else None)
            spare.append(G3)
            if pn:
                spare.append(G4)
            # the column partials of the fc2 data-gradient GEMM and of this LayerNorm are finished at the end of the block, in the
            # same launch as the attention / LayerNorm 1 partials (nothing in between reads their results)
            mlp_jobs += [(self.pg2_, R, D, self.g(pre + "norm2.weight")), (self.pb2_, R, D, self.g(pre + "norm2.bias")),
                         (self.pd2_, R, D, self.g(pre + "attn.proj.bias"))]
            # proj
            ops.gemm(ops.EPI_WGRAD, G2, a["o"], M=D, N=A, K=M, out0=self.g(pre + "attn.proj.weight"), a_mn=True, b_mn=True)
            ops.gemm(ops.EPI_STORE, G2, self.w(pre + "attn.proj.weight"), M=M, N=A, K=D, out0=dO, b_mn=True, rowscale=dp1,
                     rows_per_scale=T, bn=self.bn_nD if A == D else 0)
            # attention
            ops.attention_bwd(a["qkv"], a["o"], dO, a["lse"], g_a, dp1, dqkv, self.att_pg, self.att_pb, B, T, H,
                              self.scale)
            a_off = bm.modules[i_a]["gate_off"]
            attn_jobs = [dict(part=self.att_pg, R=B, N=A, out=self.dgate[a_off:a_off + A], div_by=g_a, accumulate=False),
                         dict(part=self.att_pb, R=B, N=3 * A, out=self.g(pre + "attn.qkv.bias"))]
            ops.gemm(ops.EPI_WGRAD, dqkv, a["x1"], M=3 * A, N=D, K=M, out0=self.g(pre + "attn.qkv.weight"), a_mn=True,
                     b_mn=True)
            G1 = spare.pop()
            ops.gemm(ops.EPI_STORE, dqkv, self.w(pre + "attn.qkv.weight"), M=M, N=D, K=3 * A, out0=G1, b_mn=True,
                     res=None if pn else G2, bn=self.bn_nD)
            if not pn:
                spare.append(G2)
            # LayerNorm 1 (+ previous block's fc2 bias grad)
            G0 = spare.pop()
            has_prev = l > 0
            ops.layernorm_bwd(G1, self.xs[l], a["mean1"], a["rstd1"], self.p(pre + "norm1.weight"), G0, self.pg_,
                              self.pb_, self.pd_ if has_prev else None,
                              self.drop_scale[2 * l - 1] if has_prev else None, T, d_valid=Dv, dres=G2 if pn else None)
            spare.append(G1)
            if pn:
                spare.append(G2)
            ln1_jobs = [(self.pg_, R, D, self.g(pre + "norm1.weight")), (self.pb_, R, D, self.g(pre + "norm1.bias"))]
            if has_prev:
                ln1_jobs.append((self.pd_, R, D, self.g(f"blocks.{l - 1}.mlp.fc2.bias")))
            ops.reduce_partials_multi(mlp_jobs + attn_jobs + ln1_jobs)
            ops.wgrad_flush()
            G = G0
            if red is not None:
                ops.wgrad_join()          # the bucket all-reduce of this block reads the finished weight gradients
                red.on_block_done(l)

        # ---- embed stage ----
        g_e = bm.gate_of(0)
        ops.embed_bwd(G, self.xs[0], g_e, self.mask if pmim else self.zero_mask, self.dconv, self.e_gx, self.e_pos, self.e_mt,
                      B, T, D)
        embed_jobs = [
            dict(part=self.e_gx, R=T, N=D, out=self.dgate[0:D], div_by=g_e, accumulate=False),
            (self.e_pos, 1, T * D, self.g("pos_embed")), (self.e_pos, 1, D, self.g("cls_token")),
            (self.e_pos[1:], L, D, self.g("patch_embed.proj.bias"))]
        if pmim and not self.decoder_frozen:                 # freeze_decoder() also freezes the mask token (vt:535-536)
            embed_jobs.append((self.e_mt, T, D, self.g("mask_token")))
        ops.reduce_partials_multi(embed_jobs)
        ops.gemm(ops.EPI_WGRAD, self.dconv, self.patches, M=D, N=768, K=ML,
                 out0=self.g("patch_embed.proj.weight").view(D, 768), a_mn=True, b_mn=True)
        # ---- bi-mask: d gate (+ FLOPs / sparsity losses) -> d score, d alpha ----
        bm.backward(self.params, self.hyper[self._wp_idx:self._wp_idx + 1], self.dgate, gs if loss_grads else 0.0, self.grads)
        ops.wgrad_flush()
        ops.wgrad_join()
        if red is not None:
            red.finish()

    # ------------------------------------------------------------------------------------------------------------
    def allreduce_grads(self):
        """Data-parallel mean of every gradient (DDP, search.py:619): a few large NCCL all-reduces over the flat arena."""
        dp.allreduce_arena(self.grads, self.world, self.pg, self._dp_bounds)

    def optimizer_step(self):
        ops.adamw(self.params, self.grads, self.adam_m, self.adam_v, self.shadow, self.hyper, self._seg_end_c,
                  zero_grad=True)
        self.step_count += 1

    def step_graphed(self, images, labels, lrs=None, target=None, update=True, noise=None):
        """One full search step replayed from a CUDA graph (forward, backward, [gradient exchange,] AdamW; the whole step is
        ~255 launches of ~10-200 us, so per-launch host work would otherwise bound it). The graph is captured on first use for
        this (images buffer, labels buffer, PMIM keep count, update flag) and re-captured when the schedule changes the keep
        count; lr, AdamW bias corrections and w_p are read from the device-side `hyper` vector, so they may change every step.
        Random draws (PMIM noise, DropPath) come from torch's graph-safe generator; `noise` (optional, [B, L]) pins the PMIM
        draw instead (copied into a persistent buffer the graph reads).
        update=False is a gradient-accumulation micro-step (engine.py:152, 169: backward on every micro-step, optimizers only
        when (data_iter_step + 1) % accum_iter == 0): its graph ends after backward, gradients keep accumulating in the arena,
        no exchange, no AdamW, and step_count does not advance. The reference all-reduces on every micro-step (no no_sync,
        SURVEY App. B-9); averaging once on the boundary gives the same mean.
        Multi-GPU: with OFB_DP_GRAPH=1 the bucket all-reduces and the update are nodes of the same graph (no host launches
        between backward and the update); otherwise they follow the replay from the host."""
        keep = int(self.L * self.keep_ratio)
        key = (images.data_ptr(), labels.data_ptr() if labels is not None else 0, keep,
               target.data_ptr() if target is not None else 0, bool(update), noise is not None)
        self._fill_hyper(lrs)
        self._hyper_up.upload(self.hyper)
        if noise is not None:
            if self._noise_in is None:
                self._noise_in = torch.empty(self.B, self.L, dtype=torch.float32, device=self.dev)
            self._noise_in.copy_(noise)
        nz = self._noise_in if noise is not None else None
        in_graph_dp = self.world > 1 and (self.dp_overlap or self.dp_in_graph)
        entry = self._graphs.get(key)
        if entry is None:
            # warm-up outside capture: first launches configure kernel attributes and load modules. Gradients accumulated by
            # earlier micro-steps must survive it.
            cur = torch.cuda.current_stream(self.dev)
            side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                saved = self.grads.clone()
                self.forward(images, labels, target=target, noise=nz)
                self.backward(exchange=self.dp_overlap)      # also brings up the NCCL communicator before capture
                if self.world > 1 and self.dp_in_graph and not self.dp_overlap:
                    self.allreduce_grads()
                self.grads.copy_(saved)
                del saved
            cur.wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            n0 = ops.LAUNCHES
            with torch.cuda.graph(graph):
                self.forward(images, labels, target=target, noise=nz)
                self.backward(exchange=self.dp_overlap and update)      # overlapped exchange: the NCCL launches are graph nodes
                if update:
                    if self.world > 1 and self.dp_in_graph and not self.dp_overlap:
                        self.allreduce_grads()
                    if self.world <= 1 or in_graph_dp:
                        ops.adamw(self.params, self.grads, self.adam_m, self.adam_v, self.shadow, self.hyper, self._seg_end_c,
                                  zero_grad=True)
            entry = (graph, ops.LAUNCHES - n0)
            self._graphs[key] = entry
        entry[0].replay()
        ops._count(entry[1])
        if update:
            if self.world > 1 and not in_graph_dp:
                self.allreduce_grads()
                ops.adamw(self.params, self.grads, self.adam_m, self.adam_v, self.shadow, self.hyper, self._seg_end_c,
                          zero_grad=True)
            self.step_count += 1
        return self.scal

    def evaluate(self, images, labels):
        """One evaluate() batch (engine.py:222-257) in the reference's eval mode of an unfinished search: returns the device
        tensor [mean cross entropy, top-1 fraction, top-5 fraction] of this batch; logits stay in self.logits."""
        self._hyper_up.begin(keep=True)[self._wp_idx] = self.w_p
        self._hyper_up.upload(self.hyper)
        return self.forward(images, labels, train=False)

    # ------------------------------------------------------------------------------------------------------------
    # prune event (vision_transformer.py:785-950; see prune.py)
    def plan_prune(self, thresh=0.2):
        """Decisions and kept-unit index sets of compress(thresh) for every searchable module, from the alphas (averaged over
        the ranks, layers.py:9-14), the switch cells and the ranks the bi-mask kernel built in the last forward."""
        from . import prune
        bm = self.bimask
        rank = bm.rank.cpu()
        plans = {}
        for i, m in enumerate(bm.modules):
            pre, H, dim = m["prefix"], m["heads"], m["dim"]
            alpha = self.p(pre + ".alpha").detach().clone()
            if self.world > 1:
                torch.distributed.all_reduce(alpha, group=self.pg)
                alpha /= self.world
            r = bm.logical(i, rank)
            widths = bm._widths[m["width_off"]:m["width_off"] + m["n_j"]]
            counts = bm._widths[m["width_off"] + m["n_j"]:m["width_off"] + m["n_j"] + (m["n_i"] if m["kind"] == 2 else 0)]
            sw = self.switches[pre].reshape(m["n_i"], m["n_j"])
            plans[pre] = prune.plan_module(pre, m["kind"], alpha.reshape(m["n_i"], m["n_j"]), sw, list(widths), list(counts),
                                           r[:, 0] // dim, r % dim, thresh)
        return plans

    def gather_pruned(self, plans):
        """Every parameter in the shape compress() would leave it in (device tensors, reference names / shapes)."""
        from . import prune
        dims = {m["prefix"]: dict(heads=m["heads"], dim=m["dim"]) for m in self.bimask.modules}
        return prune.gather_pruned(plans, self.named_parameters(), dims, self.w_p)

    def pruned_config(self, plans):
        """Constructor argument `pruned` (and the surviving switch cells) of the engine a set of plans leaves behind."""
        bm = self.bimask
        spaces, heads, hdims, hids, switches = {}, [], [], [], {}
        embed = self.Dv
        for m in bm.modules:
            pl = plans[m["prefix"]]
            n_i, n_j = pl.switch.shape
            wl = bm._widths[m["width_off"]:m["width_off"] + m["n_j"] + (m["n_i"] if m["kind"] == 2 else 0)]
            spaces[m["prefix"]] = (list(wl[:n_j]), list(wl[m["n_j"]:m["n_j"] + n_i]) if m["kind"] == 2 else [])
            switches[m["prefix"]] = pl.switch.clone()
            if m["kind"] == 0:
                embed = pl.width if pl.truncated else m["dim"]
            elif m["kind"] == 2:
                heads.append(pl.head_num if pl.truncated else m["heads"])
                hdims.append(pl.width if pl.truncated else m["dim"])
            else:
                hids.append(pl.width if pl.truncated else m["dim"])
        return dict(embed=embed, heads=heads, head_dims=hdims, hiddens=hids, spaces=spaces), switches

    def rebuild_pruned(self, plans):
        """The search engine after a TRUNCATING prune event (compress(), vision_transformer.py:785-950): a new engine on the
        sliced shapes with the gathered parameters and the sliced Adam state (optim.AdamW.update, optim.py:122-182: moments
        follow the same index gathers; the alphas of modules that executed a prune restart from zero with their own step
        counter, and so do finalised scores). Finished modules gate with their frozen score; a finished embedding search switches
        the blocks to standard pre-norm."""
        from . import prune
        pruned, switches = self.pruned_config(plans)
        eng = SearchStepEngine(self.D0, self.H, self.depth, self.B, mlp_ratio=self.hid // self.D0, num_classes=self.C,
                               img=self.img, patch=self.P, drop_path_rate=self.drop_path_rate, lr=self.lr, weight_decay=self.wd,
                               eps_ln=self.eps_ln, smoothing=self.smoothing, accum_iter=self.accum_iter,
                               warmup_epochs=self.warmup_epochs, max_ratio=self.max_ratio, min_ratio=self.min_ratio,
                               device=self.dev, switches=switches, process_group=self.pg, pruned=pruned, **self._loss_w)
        dims = {m["prefix"]: dict(heads=m["heads"], dim=m["dim"]) for m in self.bimask.modules}
        eng.load_params(prune.gather_pruned(plans, self.named_parameters(), dims, self.w_p))
        for arena_old, arena_new in ((self.adam_m, eng.adam_m), (self.adam_v, eng.adam_v)):
            state = prune.gather_pruned(plans, {k: self._unpad(k, self._view(arena_old, k)) for k in self.offsets}, dims,
                                        self.w_p, state=True)
            for k in eng.offsets:
                eng._view(arena_new, k).copy_(eng._pad(k, state[k].to(self.dev, torch.float32).reshape(eng.ref_shapes[k])))
        eng.step_count = self.step_count
        eng.w_p, eng.keep_ratio = self.w_p, self.keep_ratio
        eng._schedule_touched, eng.decoder_frozen = self._schedule_touched, self.decoder_frozen
        for pre, pl in plans.items():
            eng.alpha_restart[pre + ".alpha"] = self.step_count if pl.executed else self.alpha_restart[pre + ".alpha"]
            eng.alpha_restart[pre + ".score"] = self.step_count if pl.finalised else self.alpha_restart[pre + ".score"]
        return eng

    def prune_event(self, thresh=0.2):
        """engine.py:201-213 (three times per epoch): model.compress(thresh, optimizers). Returns (engine to continue with,
        finish_search, execute_prune): the same engine when the event slices nothing (switch-only events are applied in place),
        a rebuilt engine on the sliced shapes otherwise. The caller drops its reference to the old engine."""
        plans = self.plan_prune(thresh)
        executed = any(pl.executed for pl in plans.values())
        finished = all(pl.finished for pl in plans.values())
        if any(pl.truncated for pl in plans.values()):
            return self.rebuild_pruned(plans), finished, executed
        self.apply_prune(plans)
        return self, finished, executed

    def apply_prune(self, plans):
        """Apply a prune event IN PLACE. Only events that switch cells off without slicing any tensor qualify: new switch
        cells, alpha <- where(alive, mean alpha, 0), Adam state of those alphas restarted (optim.py:152-159), step graphs
        dropped. Truncating / finalising events change shapes: use rebuild_pruned() (or prune_event(), which picks)."""
        if any(pl.truncated for pl in plans.values()):
            raise NotImplementedError("this prune event slices tensors: rebuild_pruned(plans) builds the engine on the new shapes")
        changed = False
        for pre, pl in plans.items():
            if not pl.executed:
                continue
            changed = True
            self.switches[pre] = pl.switch.clone()
            self._finish_cache = None
            name = pre + ".alpha"
            self.p(name).copy_(pl.alpha.to(self.dev).reshape(self.shapes[name]))
            self._view(self.adam_m, name).zero_()
            self._view(self.adam_v, name).zero_()
            self.alpha_restart[name] = self.step_count
        if changed:
            bm = self.bimask
            sw_bytes = []
            for m in bm.modules:
                sw_bytes.extend(int(x) for x in self.switches[m["prefix"]].reshape(-1).tolist())
            bm._sw_bytes = sw_bytes
            bm.switches_dev.copy_(torch.tensor(sw_bytes, dtype=torch.uint8))
            self.sync_shadow()
            self.release_graphs()
        return changed

    # ------------------------------------------------------------------------------------------------------------
    # checkpoint / resume in a state-dict format with shape metadata. The reference pickles the whole model object
    # plus three optimizer state_dicts (search.py:671-740) and unpickles it in resume() (search.py:302-372) because compress()
    # changes shapes; here the shapes travel as metadata next to the tensors, which keep the reference's names and shapes.
    def state_dict(self):
        named = lambda arena: {k: self._unpad(k, self._view(arena, k)).detach().cpu().clone() for k in self.offsets}
        pruned = None if self.spaces is None else dict(embed=self.Dv, heads=list(self.heads), head_dims=list(self.hdims),
                                                       hiddens=list(self.hids),
                                                       spaces={k: (list(v[0]), list(v[1])) for k, v in self.spaces.items()})
        return dict(
            format="ofb_b200.search/1",
            config=dict(embed_dim=self.D0, num_heads=self.H, depth=self.depth, mlp_ratio=self.hid // self.D0, num_classes=self.C,
                        img=self.img, patch=self.P, drop_path_rate=self.drop_path_rate, lr=self.lr, weight_decay=self.wd,
                        eps_ln=self.eps_ln, smoothing=self.smoothing, accum_iter=self.accum_iter, warmup_epochs=self.warmup_epochs,
                        max_ratio=self.max_ratio, min_ratio=self.min_ratio, **self._loss_w),
            pruned=pruned, switches={k: v.clone().cpu() for k, v in self.switches.items()},
            params=named(self.params), adam_m=named(self.adam_m), adam_v=named(self.adam_v),
            step_count=self.step_count, alpha_restart=dict(self.alpha_restart), w_p=self.w_p, keep_ratio=self.keep_ratio,
            schedule_touched=self._schedule_touched, decoder_frozen=self.decoder_frozen)

    @classmethod
    def from_state_dict(cls, sd, batch, device="cuda", process_group=None):
        """Rebuild the engine a state_dict() was taken from (any batch size): same shapes (pruned or not), switch cells, parameters,
        Adam moments with their per-tensor restart steps, schedule and post-search state."""
        if sd.get("format") != "ofb_b200.search/1":
            raise ValueError(f"not a SearchStepEngine checkpoint: format {sd.get('format')!r}")
        eng = cls(batch=batch, device=device, process_group=process_group, switches={k: v.clone() for k, v in sd["switches"].items()},
                  pruned=sd["pruned"], **sd["config"])
        if set(sd["params"]) != set(eng.offsets):
            raise ValueError("checkpoint tensors do not match the engine's parameter set")
        eng.load_params(sd["params"])
        for arena, key in ((eng.adam_m, "adam_m"), (eng.adam_v, "adam_v")):
            for k in eng.offsets:
                eng._view(arena, k).copy_(eng._pad(k, sd[key][k].to(eng.dev, torch.float32).reshape(eng.ref_shapes[k])))
        eng.step_count = int(sd["step_count"])
        eng.alpha_restart.update(sd["alpha_restart"])
        eng.w_p, eng.keep_ratio = float(sd["w_p"]), float(sd["keep_ratio"])
        eng._schedule_touched, eng.decoder_frozen = bool(sd["schedule_touched"]), bool(sd["decoder_frozen"])
        return eng

    def release_graphs(self):
        """Drop every captured step graph (required before the process group is destroyed when the graphs hold NCCL
        launches, and after a prune event changes shapes)."""
        self._graphs.clear()

    def step(self, images, labels, noise=None, drop_u=None, update=True, lrs=None, target=None, mix=None):
        """One full search step; returns the device tensor scal = [base, arch, decoder, total, w_dec, ...].
        Post-search phase (after enter_post_search): target = Mixup soft targets [B, C]; mix = the MixParams when `images` is the
        unmixed batch (blend fused into the im2col)."""
        self._fill_hyper(lrs)
        self._hyper_up.upload(self.hyper)
        self.forward(images, labels, noise, drop_u, target=target, mix=mix)
        self.backward(exchange=update and self.dp_overlap)
        if update:
            if not self.dp_overlap:
                self.allreduce_grads()
            self.optimizer_step()
        return self.scal
