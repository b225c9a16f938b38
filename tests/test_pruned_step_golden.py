"""Search step AFTER a truncating prune event (the post-prune search shapes): the generalised CPU oracle
(forward_step(shape=...)) fed with the planner's gathers against the UNMODIFIED reference's compress() + search step
(tests/golden/pruned_step/*.npz, oracle/make_golden_pruned_step.py), and the GPU engine rebuilt on the pruned shapes against
both (gpu marker)."""
import glob
import os

import numpy as np
import pytest
import torch

from fixtures import make_inputs, make_params, pruned_shape_from_plans, summarize
from make_golden_pruned_step import plan_on_cpu, script
from ofb_oracle import ModelCfg, default_switches, train_step, w_p_schedule

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pruned_step", "*.npz")))
IDS = [os.path.basename(p)[:-4] for p in GOLD]


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def _case(path):
    g = np.load(path)
    D, H, depth, B = (int(x) for x in g["cfg"])
    cfg = ModelCfg(embed_dim=D, num_heads=H, depth=depth)
    c = dict(D=D, H=H, depth=depth, offset=int(g["offset"]), mixed=bool(g["mixed"]))
    P0 = script(make_params(cfg, seed=0), c)
    inp = make_inputs(cfg, B, seed=1, epoch_frac=float(g["epoch_frac"]), drop_path_rate=float(g["dpr"]))
    return g, cfg, P0, inp


@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_oracle_pruned_step_matches_reference_golden(path):
    import ofb_b200  # noqa: F401
    from ofb_b200 import prune
    g, cfg, P0, inp = _case(path)
    plans, dims = plan_on_cpu(cfg, P0, default_switches(cfg))
    Pp = prune.gather_pruned(plans, {k: v for k, v in P0.items() if k != "alpha_patch"}, dims, w_p_schedule(float(g["epoch_frac"])))
    Pp["alpha_patch"] = P0["alpha_patch"]
    shape = pruned_shape_from_plans(cfg, plans)
    assert (shape.embed, shape.heads, shape.head_dims, shape.hiddens) == (int(g["embed"]), g["heads"].tolist(),
                                                                          g["head_dims"].tolist(), g["hiddens"].tolist())
    out, grads = train_step(Pp, {}, inp, cfg, lr=float(g["lr"]), step=1, switches={k: pl.switch for k, pl in plans.items()},
                            shape=shape)
    tol = 1e-4
    assert _rel(out.logits.detach().numpy(), g["logits"]) < tol
    for name, val in (("loss_base", out.loss_base), ("loss_arch", out.loss_arch), ("loss_decoder", out.loss_decoder),
                      ("loss_total", out.loss_total)):
        assert _rel(val.detach().numpy(), g[name]) < tol, name
    assert _rel(float(out.loss_terms["flops_searched"]), g["flops"][1]) < tol
    n = 0
    for key in g.files:
        if key.startswith("gsum:"):
            assert _rel(summarize(grads[key[5:]]).numpy(), g[key]) < tol, key
            n += 1
    assert n > 20


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_engine_pruned_step_matches_reference_golden(cuda_dev, path):
    """SearchStepEngine -> plan_prune -> rebuild_pruned (new engine on the truncated shapes, zero-padded layout) -> one
    search step, against the generalised oracle (every gradient) and the reference's own outputs (golden)."""
    from ofb_b200.engine import SearchStepEngine
    from step_compare import BF16_TOL, FP32_TOL, GRAD_MAX_TOL, LOSS_TOL, elementwise_bound_applies, rel, rel_l2
    g, cfg, P0, inp = _case(path)
    B, depth = inp.images.shape[0], cfg.depth
    dpr, lr, ef = float(g["dpr"]), float(g["lr"]), float(g["epoch_frac"])
    eng0 = SearchStepEngine(cfg.embed_dim, cfg.num_heads, depth, B, drop_path_rate=dpr, lr=lr)
    eng0.load_params(P0)
    eng0.set_schedule(ef)
    img, lab, noise = inp.images.cuda(), inp.labels.cuda(), inp.noise.cuda()
    drop_u = ((inp.drop_scale > 0).float().reshape(depth * 2, B) * 0.999).cuda()
    eng0.step(img, lab, noise=noise, drop_u=drop_u, update=False)           # ranks for the plan
    eng0.grads.zero_()
    plans = eng0.plan_prune(0.2)
    assert any(pl.truncated for pl in plans.values())
    assert any(pl.finalised for pl in plans.values()) == bool(g["mixed"])
    eng = eng0.rebuild_pruned(plans)
    assert (eng.Dv, eng.heads, eng.hdims, eng.hids) == (int(g["embed"]), g["heads"].tolist(), g["head_dims"].tolist(),
                                                        g["hiddens"].tolist())
    scal = eng.step(img, lab, noise=noise, drop_u=drop_u, update=False)
    torch.cuda.synchronize()
    scal = scal.cpu()
    assert eng.padding_is_clean()
    # --- against the reference's own numbers ---
    assert _rel(eng.logits.cpu().numpy(), g["logits"]) < BF16_TOL
    assert _rel(float(scal[0]), g["loss_base"]) < LOSS_TOL
    assert _rel(float(scal[1]), g["loss_arch"]) < 1e-3
    assert _rel(float(scal[2]), g["loss_decoder"]) < LOSS_TOL
    assert _rel(float(scal[3]), g["loss_total"]) < LOSS_TOL
    assert _rel(float(eng.bimask.arch[5]), g["flops"][1]) < 1e-4 and _rel(float(eng.bimask.arch[6]), g["flops"][0]) < 1e-4
    # --- every gradient against the oracle on the same pruned parameters ---
    Pp = {k: v.detach().cpu().clone() for k, v in eng.named_parameters().items()}
    Pp["alpha_patch"] = P0["alpha_patch"]
    shape = pruned_shape_from_plans(cfg, plans)
    out, grads = train_step(Pp, {}, inp, cfg, lr=lr, step=1, switches={k: pl.switch for k, pl in plans.items()}, shape=shape)
    got = eng.named_grads()
    worst_l2, worst_max = ("", 0.0), ("", 0.0)
    for k, gr in grads.items():
        if gr is None:
            continue
        e2, em = rel_l2(got[k], gr), rel(got[k], gr)
        worst_l2 = max(worst_l2, (k, e2), key=lambda kv: kv[1])
        if elementwise_bound_applies(k, B):
            worst_max = max(worst_max, (k, em), key=lambda kv: kv[1])
    print("worst L2", worst_l2, "worst max", worst_max)
    # fixed bound of step_compare.py for EVERY gradient tensor (decoder, alphas and scores included)
    assert worst_l2[1] < BF16_TOL and worst_max[1] < GRAD_MAX_TOL
    for i, m in enumerate(eng.bimask.modules):
        assert rel(eng.bimask.logical(i, eng.bimask.gate).reshape(-1), out.gates[m["prefix"]].reshape(-1)) < FP32_TOL
    # the update keeps the padding clean and the engine can keep stepping (graph replay included)
    eng.grads.zero_()
    eng.step(img, lab, noise=noise, drop_u=drop_u)
    eng.step_graphed(img, lab)
    torch.cuda.synchronize()
    assert eng.padding_is_clean() and torch.isfinite(eng.scal).all()


@pytest.mark.gpu
def test_consecutive_prune_events(cuda_dev):
    """A trajectory engine.py:201-213 style: steps, prune event, steps, second prune event ON THE ALREADY PRUNED ENGINE. The
    second event's plan (alphas, switches, ranks read from the pruned engine's strided tables) must equal the CPU planner's
    on the same parameters, and the twice-rebuilt engine must still match the oracle on its shapes."""
    import ofb_b200  # noqa: F401
    from fixtures import search_modules
    from ofb_b200 import prune
    from ofb_b200.engine import SearchStepEngine
    from ofb_oracle import PrunedShape, _desc_rank, forward_step
    from step_compare import BF16_TOL, FP32_TOL, LOSS_TOL, rel
    cfg = ModelCfg(embed_dim=192, num_heads=3, depth=2)
    c = dict(D=192, H=3, depth=2, offset=0, mixed=False)
    P0 = script(make_params(cfg, seed=0), c)
    B = 2
    inp = make_inputs(cfg, B, seed=1, epoch_frac=6.0, drop_path_rate=0.0)
    img, lab, noise = inp.images.cuda(), inp.labels.cuda(), inp.noise.cuda()
    eng = SearchStepEngine(192, 3, 2, B, drop_path_rate=0.0, lr=1e-3)
    eng.load_params(P0)
    eng.set_schedule(6.0)
    for _ in range(2):
        eng.step(img, lab, noise=noise)
    eng, finished, executed = eng.prune_event(0.2)
    assert executed and not finished and eng.spaces is not None
    for _ in range(2):
        eng.step(img, lab, noise=noise)
    # second event: kill the trailing columns of every remaining alpha table on the pruned engine
    for name in eng.alpha_names:
        a = eng.p(name)
        if a.shape[-1] > 2:
            a[..., -1:] = -9.0
    eng.step(img, lab, noise=noise, update=False)
    eng.grads.zero_()
    plans = eng.plan_prune(0.2)
    # CPU planner on the same (logical) parameters of the pruned engine
    Pn = {k: v.detach().cpu().clone() for k, v in eng.named_parameters().items()}
    for i, m in enumerate(eng.bimask.modules):
        pre, H, dim = m["prefix"], m["heads"], m["dim"]
        score = Pn[pre + ".score"].reshape(H, dim)
        hr = _desc_rank(torch.sigmoid(score).sum(-1)) if H > 1 else torch.zeros(1, dtype=torch.long)
        widths, counts = eng.spaces[pre]
        ref_pl = prune.plan_module(pre, m["kind"], Pn[pre + ".alpha"], eng.switches[pre], list(widths), list(counts), hr,
                                   _desc_rank(score), 0.2)
        got = plans[pre]
        assert torch.equal(got.switch, ref_pl.switch), pre
        assert (got.truncated, got.finalised, got.width, got.head_num) == (ref_pl.truncated, ref_pl.finalised, ref_pl.width,
                                                                          ref_pl.head_num), pre
        if got.truncated:
            assert torch.equal(got.channel_index, ref_pl.channel_index) and torch.equal(got.head_index, ref_pl.head_index), pre
    assert any(pl.truncated for pl in plans.values())
    eng2 = eng.rebuild_pruned(plans)
    scal = eng2.step(img, lab, noise=noise, update=False)
    torch.cuda.synchronize()
    assert eng2.padding_is_clean()
    P2 = {k: v.detach().cpu().clone() for k, v in eng2.named_parameters().items()}
    P2["alpha_patch"] = P0["alpha_patch"]
    shape = PrunedShape(embed=eng2.Dv, heads=eng2.heads, head_dims=eng2.hdims, hiddens=eng2.hids, spaces=eng2.spaces)
    inp.w_p, inp.keep_ratio = eng2.w_p, eng2.keep_ratio
    out = forward_step(P2, inp, cfg, {k: v for k, v in eng2.switches.items()}, shape)
    assert rel(eng2.logits, out.logits) < BF16_TOL
    assert rel(scal[0], out.loss_base) < LOSS_TOL and rel(scal[1], out.loss_arch) < 10 * FP32_TOL
    assert rel(scal[3], out.loss_total) < LOSS_TOL
