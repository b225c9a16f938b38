"""Hand-off from the search to the finetune / fused-inference model: prune plan -> exact gathers ->
fuse (gates folded into the weights) -> plain pruned ViT, against the UNMODIFIED reference's compress() + fuse() + eval
forward (tests/golden/fuse/*.npz, oracle/make_golden_fuse.py). CPU: planner with torch ranks, logits from the finetune
oracle. GPU (gpu marker): SearchStepEngine -> plan_prune -> gather_pruned -> fuse_params -> FinetuneStepEngine.evaluate, every
tensor staying on the device."""
import glob
import os

import numpy as np
import pytest
import torch

from fixtures import make_inputs, make_params, summarize
from make_golden_fuse import script_single_survivor
from ofb_oracle import (ModelCfg, _desc_rank, default_switches, embed_widths, head_channel_widths, head_counts, hidden_widths,
                        w_p_schedule)

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fuse", "*.npz")))
IDS = [os.path.basename(p)[:-4] for p in GOLD]


def _case(path):
    g = np.load(path)
    D, H, depth, B = (int(x) for x in g["cfg"])
    cfg = ModelCfg(embed_dim=D, num_heads=H, depth=depth)
    P = script_single_survivor(make_params(cfg, seed=0))
    inp = make_inputs(cfg, B, seed=1, epoch_frac=float(g["epoch_frac"]), drop_path_rate=0.0, keep_ratio=1.0)
    return g, cfg, P, inp


def _check_tensors(g, fused):
    n = 0
    for key in g.files:
        if key.startswith("shape:"):
            assert tuple(fused[key[6:]].shape) == tuple(int(x) for x in g[key]), key
        elif key.startswith("sum:"):
            got = summarize(fused[key[4:]].cpu()).numpy()
            # products with the finalised scores (w_p sigmoid(s) + 1 - w_p, 1 ulp): fp32 tolerance
            assert np.allclose(got, g[key], rtol=2e-6, atol=1e-9), key
            n += 1
    assert n == len(fused)


@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_fuse_handoff_matches_reference(path):
    import ofb_b200  # noqa: F401
    from ft_oracle import SubnetCfg, ft_forward
    from ofb_b200 import prune
    g, cfg, P, inp = _case(path)
    w_p = w_p_schedule(float(g["epoch_frac"]))
    sw = default_switches(cfg)
    mods = [("patch_embed", 0, 1, cfg.embed_dim, embed_widths(cfg.embed_dim), [])]
    for l in range(cfg.depth):
        mods.append((f"blocks.{l}.attn", 2, cfg.num_heads, cfg.head_dim, head_channel_widths(cfg.head_dim), head_counts(cfg.num_heads)))
        mods.append((f"blocks.{l}.mlp", 1, 1, cfg.hidden, hidden_widths(cfg.hidden), []))
    plans, dims = {}, {}
    for prefix, kind, H, dim, widths, heads in mods:
        score = P[prefix + ".score"].reshape(H, dim)
        hr = _desc_rank(torch.sigmoid(score).sum(-1)) if H > 1 else torch.zeros(1, dtype=torch.long)
        plans[prefix] = prune.plan_module(prefix, kind, P[prefix + ".alpha"], sw[prefix], widths, heads, hr, _desc_rank(score), 0.2)
        dims[prefix] = dict(heads=H, dim=dim)
    D2, heads, head_dims, hiddens = prune.subnet_dims(plans, cfg.depth)
    assert (D2, heads, head_dims, hiddens) == (int(g["embed"]), g["heads"].tolist(), g["head_dims"].tolist(), g["hiddens"].tolist())
    pruned = prune.gather_pruned(plans, {k: v for k, v in P.items() if k != "alpha_patch"}, dims, w_p)
    fused = prune.fuse_params(pruned, list(plans))
    _check_tensors(g, fused)
    sub = SubnetCfg(embed_dim=D2, heads=heads, head_dims=head_dims, hiddens=hiddens, scale=cfg.head_dim ** -0.5)
    logits = ft_forward(fused, inp.images, sub)
    ref = torch.from_numpy(g["logits"])
    assert float((logits - ref).abs().max() / ref.abs().max()) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_engine_fuse_handoff_matches_reference(cuda_dev, path):
    from ofb_b200 import prune
    from ofb_b200.engine import SearchStepEngine
    from ofb_b200.finetune_engine import FinetuneStepEngine
    from step_compare import BF16_TOL
    g, cfg, P, inp = _case(path)
    B = inp.images.shape[0]
    eng = SearchStepEngine(cfg.embed_dim, cfg.num_heads, cfg.depth, B, drop_path_rate=0.0)
    eng.load_params(P)
    eng.set_schedule(float(g["epoch_frac"]))
    img, lab = inp.images.cuda(), inp.labels.cuda()
    eng.step(img, lab, update=False)                       # builds the ranks the plan reads
    plans = eng.plan_prune(0.2)
    fused = prune.fuse_params(eng.gather_pruned(plans), list(plans))
    assert all(v.is_cuda for v in fused.values())
    _check_tensors(g, fused)
    D2, heads, head_dims, hiddens = prune.subnet_dims(plans, cfg.depth)
    ft = FinetuneStepEngine(D2, heads, head_dims, hiddens, B, attn_scale=eng.scale)
    ft.load_params(fused)
    out = ft.evaluate(img, lab)
    torch.cuda.synchronize()
    ref = torch.from_numpy(g["logits"])
    assert float((ft.logits.cpu() - ref).abs().max() / ref.abs().max()) < BF16_TOL
    assert ft.padding_is_clean() and float(out[0]) > 0
