"""Prune event (SURVEY §8f rank 1, §4 iii): the planner / gather of ofb_b200.prune against what the UNMODIFIED reference's
compress() leaves behind on scripted alpha trajectories (tests/golden/prune/*.npz, oracle/make_golden_prune.py): switch
cells, alphas, finish / execute flags, scores, shapes of every tensor and - the gathers being exact - bit-identical
fingerprints of every pruned tensor. CPU: ranks from torch; GPU (gpu marker): ranks from the bi-mask forward kernel."""
import glob
import os

import numpy as np
import pytest
import torch

from fixtures import make_params, summarize
from make_golden_prune import script_alphas
from ofb_oracle import (ModelCfg, _desc_rank, default_switches, embed_widths, head_channel_widths, head_counts, hidden_widths,
                        w_p_schedule)

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prune", "*.npz")))
IDS = [os.path.basename(p)[:-4] for p in GOLD]


def _modules(cfg):
    mods = [("patch_embed", 0, 1, cfg.embed_dim, embed_widths(cfg.embed_dim), [])]
    for l in range(cfg.depth):
        mods.append((f"blocks.{l}.attn", 2, cfg.num_heads, cfg.head_dim, head_channel_widths(cfg.head_dim),
                     head_counts(cfg.num_heads)))
        mods.append((f"blocks.{l}.mlp", 1, 1, cfg.hidden, hidden_widths(cfg.hidden), []))
    return mods


def _check(g, cfg, P, plans, pruned):
    for prefix, kind, H, dim, widths, heads in _modules(cfg):
        pl = plans[prefix]
        assert np.array_equal(pl.switch.numpy(), g["switch:" + prefix]), prefix
        fin, ex = (bool(x) for x in g["state:" + prefix])
        assert (pl.finished, pl.executed) == (fin, ex), prefix
    n = 0
    for key in g.files:
        if key.startswith("shape:"):
            k = key[6:]
            if k == "alpha_patch":
                continue
            assert tuple(pruned[k].shape) == tuple(int(x) for x in g[key]), k
        elif key.startswith("full:") and key != "full:alpha_patch":
            k = key[5:]
            if k.endswith(".score"):       # gathered scores are exact; finalised ones are w_p sigmoid(s) + (1 - w_p): 1 ulp
                assert np.allclose(pruned[k].cpu().numpy(), g[key], rtol=1e-6, atol=0), k
            else:
                assert np.array_equal(pruned[k].cpu().numpy(), g[key]), k      # alphas: bit-exact
        elif key.startswith("sum:"):
            k = key[4:]
            got = summarize(pruned[k].cpu()).numpy()
            assert np.array_equal(got[3:], g[key][3:]), k                         # pure gathers: the sampled entries are bit-exact
            assert np.allclose(got[:3], g[key][:3], rtol=1e-12, atol=0), k        # float64 sums: reduction order varies with host threads
            n += 1
    assert n > 10


def _case(path):
    g = np.load(path)
    D, H, depth = (int(x) for x in g["cfg"])
    cfg = ModelCfg(embed_dim=D, num_heads=H, depth=depth)
    P = script_alphas(make_params(cfg, seed=0), cfg, offset=int(g["offset"]))
    return g, cfg, P


@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_prune_plan_matches_reference_compress(path):
    import ofb_b200  # noqa: F401
    from ofb_b200 import prune
    g, cfg, P = _case(path)
    w_p = w_p_schedule(float(g["epoch_frac"]))
    sw = default_switches(cfg)
    plans, dims = {}, {}
    for prefix, kind, H, dim, widths, heads in _modules(cfg):
        score = P[prefix + ".score"].reshape(H, dim)
        head_rank = _desc_rank(torch.sigmoid(score).sum(-1)) if H > 1 else torch.zeros(1, dtype=torch.long)
        plans[prefix] = prune.plan_module(prefix, kind, P[prefix + ".alpha"], sw[prefix], widths, heads, head_rank,
                                          _desc_rank(score), 0.2)
        dims[prefix] = dict(heads=H, dim=dim)
    pruned = prune.gather_pruned(plans, {k: v for k, v in P.items() if k != "alpha_patch"}, dims, w_p)
    _check(g, cfg, P, plans, pruned)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_engine_prune_plan_matches_reference_compress(cuda_dev, path):
    """Same, with alphas / switches / ranks read from the engine after a forward (ranks built by the bi-mask kernel) and the
    gathers done on the device tensors of the parameter arena."""
    from ofb_b200.engine import SearchStepEngine
    g, cfg, P = _case(path)
    eng = SearchStepEngine(cfg.embed_dim, cfg.num_heads, cfg.depth, 2, drop_path_rate=0.1)
    eng.load_params(P)
    eng.set_schedule(float(g["epoch_frac"]))
    gen = torch.Generator().manual_seed(0)
    eng.step(torch.randn(2, 3, 224, 224, generator=gen).cuda(), torch.randint(0, 1000, (2,), generator=gen).cuda(), update=False)
    plans = eng.plan_prune(0.2)
    pruned = eng.gather_pruned(plans)
    _check(g, cfg, P, plans, pruned)
    # apply_prune() is the in-place path (switch-only events); truncating events change shapes and go through
    # rebuild_pruned() / prune_event() (tests/test_pruned_step_golden.py), so apply_prune() refuses them
    if any(pl.truncated for pl in plans.values()):
        with pytest.raises(NotImplementedError):
            eng.apply_prune(plans)


@pytest.mark.gpu
def test_switch_only_prune_event_applied_in_place(cuda_dev):
    """A prune event that only switches cells off: the engine applies it in place, the next step uses the new cells (gates and
    architecture loss against the oracle with the same switches) and the touched alphas restart their Adam state - their first
    update after the event is a first Adam step, |delta| = lr (optim.py:152-159 resets step to 0)."""
    from fixtures import make_inputs
    from ofb_b200.engine import SearchStepEngine
    from ofb_oracle import forward_step
    from step_compare import FP32_TOL, rel
    cfg = ModelCfg(embed_dim=192, num_heads=3, depth=2)
    P = make_params(cfg, seed=0)
    g = torch.Generator().manual_seed(11)
    for k in sorted(k for k in P if k.endswith(".alpha")):
        a = torch.rand(P[k].shape, generator=g)
        a.view(-1)[1] = -9.0                       # an interior cell dies: no row / column empties, nothing is sliced
        P[k] = a
    B, lr = 2, 1e-3
    eng = SearchStepEngine(192, 3, 2, B, drop_path_rate=0.0, lr=lr)
    eng.load_params(P)
    eng.set_schedule(4.0)
    inp = make_inputs(cfg, B, seed=1, epoch_frac=4.0, drop_path_rate=0.0)
    img, lab = inp.images.cuda(), inp.labels.cuda()
    for _ in range(3):                             # a few ordinary steps first: Adam state and step counter are warm
        eng.step(img, lab, noise=inp.noise.cuda())
    plans = eng.plan_prune(0.2)
    assert all(pl.executed and not pl.truncated for pl in plans.values())
    assert eng.apply_prune(plans)
    sw = {k: pl.switch.clone() for k, pl in plans.items()}
    assert all(int(s.sum()) == s.numel() - 1 for s in sw.values())
    # next step: gates / architecture loss with the new cells
    eng.step(img, lab, noise=inp.noise.cuda(), update=False)
    torch.cuda.synchronize()
    Pn = {k: v.detach().cpu().clone() for k, v in eng.named_parameters().items()}
    Pn["alpha_patch"] = P["alpha_patch"]
    for k, pl in plans.items():
        assert float(Pn[k + ".alpha"].reshape(-1)[1]) == 0.0          # dead cells carry alpha 0 (layers.py:584)
    out = forward_step(Pn, inp, cfg, sw)
    assert rel(eng.scal[1], out.loss_arch) < 10 * FP32_TOL
    for i, m in enumerate(eng.bimask.modules):
        assert rel(eng.bimask.gate_of(i), out.gates[m["prefix"]].reshape(-1)) < FP32_TOL
    # restarted Adam state: first update of every alive alpha cell has magnitude lr (+ the decoupled decay lr*wd*|alpha|)
    before = {k: Pn[k + ".alpha"].clone() for k in plans}
    eng.grads.zero_()
    eng.step(img, lab, noise=inp.noise.cuda())
    torch.cuda.synchronize()
    for k, pl in plans.items():
        after = eng.p(k + ".alpha").detach().cpu().reshape(pl.switch.shape)
        delta = (after - before[k].reshape(pl.switch.shape) * (1 - lr * 1e-3)).abs()[pl.switch]
        assert float((delta - lr).abs().max()) < 2e-2 * lr, (k, delta)
