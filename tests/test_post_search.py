"""POST-SEARCH phase of engine.search_one_epoch: after the prune event that finalises every module
(finish_search) the criterion is the base loss alone and the architecture optimizer is gone (p1: rest of that epoch, PMIM +
decoder still on); from the next epoch on masking is off, mask token + decoder are frozen and the batch goes through Mixup /
CutMix with timm SoftTargetCrossEntropy (p2, search.py:641-656).
Goldens: tests/golden/post/*.npz from the UNMODIFIED reference (oracle/make_golden_post.py). CPU: oracle vs goldens, Mixup host
logic. GPU (gpu marker): Mixup kernels bit-exact vs the restatement, SearchStepEngine in both sub-phases vs goldens + oracle."""
import glob
import os

import numpy as np
import pytest
import torch

from fixtures import make_params, summarize
from make_golden_fuse import script_single_survivor
from make_golden_post import FROZEN_P2, case_inputs, oracle_post
from ofb_oracle import ModelCfg, mixup_batch, mixup_draw, mixup_target

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "post", "*.npz")))
IDS = [os.path.basename(p)[:-4] for p in GOLD]


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def _case(path):
    g = np.load(path)
    D, H, depth, B = (int(x) for x in g["cfg"])
    cfg = ModelCfg(embed_dim=D, num_heads=H, depth=depth)
    phase = int(g["phase"])
    box = tuple(int(v) for v in g["box"])
    c = dict(D=D, H=H, depth=depth, B=B, phase=phase, compress_at=float(g["compress_at"]), epoch_frac=float(g["epoch_frac"]),
             dpr=float(g["dpr"]), lr=float(g["lr"]), lam=float(g["lam"]), box=box if any(box) else None)
    P0 = script_single_survivor(make_params(cfg, seed=0))
    inp = case_inputs(cfg, c)
    return g, cfg, c, P0, inp


def test_goldens_present():
    assert len(GOLD) >= 3 and {int(np.load(p)["phase"]) for p in GOLD} == {1, 2}


@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_oracle_post_search_matches_reference_golden(path):
    g, cfg, c, P0, inp = _case(path)
    Pp, P1, out, grads, shape, sw, plans = oracle_post(cfg, P0, inp, c)
    assert (shape.embed, shape.heads, shape.head_dims, shape.hiddens) == (int(g["embed"]), g["heads"].tolist(),
                                                                          g["head_dims"].tolist(), g["hiddens"].tolist())
    tol = 1e-4
    assert _rel(out.logits.detach().numpy(), g["logits"]) < tol
    assert _rel(float(out.loss_base), g["loss_base"]) < tol and _rel(float(out.loss_total), g["loss_total"]) < tol
    if c["phase"] == 1:
        assert _rel(float(out.loss_decoder), g["loss_decoder"]) < tol
        assert abs(float(out.loss_total) - 2 * float(out.loss_base)) < 1e-5      # base + (base/dec).detach() * dec, no arch term
    else:
        assert float(out.loss_decoder) == 0.0 and float(out.loss_total) == float(out.loss_base)
    n = 0
    for key in g.files:
        if key.startswith("gsum:"):
            assert _rel(summarize(grads[key[5:]]).numpy(), g[key]) < tol, key
            n += 1
        elif key.startswith("nograd:"):
            k = key[7:]
            assert grads.get(k) is None or float(grads[k].abs().max()) == 0.0, k
            assert torch.equal(P1[k], Pp[k]), k                                  # no gradient -> no update
    assert n > 30
    if c["phase"] == 2:
        assert all(("nograd:" + k) in g.files for k in FROZEN_P2 if k != "alpha_patch")


def test_mixup_host_draw_matches_restatement():
    """Product host logic (ofb_b200.mixup.Mixup.draw) vs the oracle's restatement of timm's batch-mode draw on the same numpy
    stream, plus the invariants of the algorithm itself."""
    import ofb_b200  # noqa: F401
    from ofb_b200.mixup import Mixup
    for kw in (dict(), dict(mixup_alpha=0.8, cutmix_alpha=0.0), dict(mixup_alpha=0.0, cutmix_alpha=1.0), dict(prob=0.5)):
        a, b = np.random.RandomState(11), np.random.RandomState(11)
        mx = Mixup(rng=a, **kw)
        seen_box = seen_blend = seen_id = False
        for _ in range(200):
            mp = mx.draw((224, 224))
            lam, box = mixup_draw(b, (224, 224), **{k: v for k, v in kw.items()})
            assert mp.lam == lam and mp.box == box
            assert 0.0 <= mp.lam <= 1.0
            if mp.box is not None:
                yl, yh, xl, xh = mp.box
                assert 0 <= yl <= yh <= 224 and 0 <= xl <= xh <= 224
                assert abs(mp.lam - (1 - (yh - yl) * (xh - xl) / 224. ** 2)) < 1e-12
                seen_box = True
            elif mp.identity:
                seen_id = True
            else:
                seen_blend = True
        assert seen_box == (kw.get("cutmix_alpha", 1.0) > 0) and seen_blend == (kw.get("mixup_alpha", 0.8) > 0)
        assert seen_id == (kw.get("prob", 1.0) < 1.0)


def test_mixup_target_rows_sum_to_one():
    y = torch.tensor([3, 7, 7, 1])
    t = mixup_target(y, 10, 0.3, 0.1)
    assert torch.allclose(t.sum(-1), torch.ones(4), atol=1e-6)
    assert abs(float(t[0, 3]) - (0.3 * 0.91 + 0.7 * 0.01)) < 1e-7 and abs(float(t[1, 7]) - 0.91) < 1e-7


# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("B,lam,box", [(4, 0.37, None), (6, 0.8125, (40, 152, 96, 224)), (5, 0.5, None), (3, 0.9, (0, 224, 0, 17)),
                                       (2, 0.999, (100, 100, 30, 60))])
def test_mixup_kernels_bit_exact(cuda_dev, B, lam, box):
    import ofb_b200  # noqa: F401
    from ofb_b200 import ops
    g = torch.Generator().manual_seed(B)
    x = torch.randn(B, 3, 224, 224, generator=g)
    y = torch.randint(0, 1000, (B,), generator=g)
    ref_x, ref_t = mixup_batch(x, y, lam, box, 1000, 0.1)
    xd, yd = x.cuda(), y.cuda()
    out = torch.empty_like(xd)
    ops.mixup_batch(xd, out, lam, box)
    assert torch.equal(out.cpu(), ref_x)                       # out of place
    tgt = torch.empty(B, 1000, device="cuda")
    ops.mixup_target(yd, tgt, lam, 0.1)
    assert torch.equal(tgt.cpu(), ref_t)
    # fused into the im2col == patchify of the materialised mix
    p_ref, p_fused = (torch.empty(B * 196, 768, dtype=torch.bfloat16, device="cuda") for _ in range(2))
    ops.patchify(out, p_ref)
    ops.patchify_mixup(xd, p_fused, lam, box)
    assert torch.equal(p_ref, p_fused)
    inplace = xd.clone()
    ops.mixup_batch(inplace, inplace, lam, box)                # in place, as timm does
    assert torch.equal(inplace, out)


@pytest.mark.gpu
def test_mixup_module_end_to_end(cuda_dev):
    import ofb_b200  # noqa: F401
    from ofb_b200.mixup import Mixup
    mx = Mixup(rng=np.random.RandomState(5))
    ref_rng = np.random.RandomState(5)
    g = torch.Generator().manual_seed(0)
    for _ in range(4):
        x = torch.randn(4, 3, 224, 224, generator=g)
        y = torch.randint(0, 1000, (4,), generator=g)
        lam, box = mixup_draw(ref_rng)
        ref_x, ref_t = mixup_batch(x, y, lam, box)
        xd, tgt = x.cuda(), torch.empty(4, 1000, device="cuda")
        mp = mx(xd, y.cuda(), tgt)
        assert (mp.lam, mp.box) == (lam, box)
        assert torch.equal(xd.cpu(), ref_x) and torch.equal(tgt.cpu(), ref_t)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_engine_post_search_matches_reference_golden(cuda_dev, path):
    """SearchStepEngine: steps -> prune_event() that finalises everything -> [enter_post_search()] -> one post-search step,
    against the reference's own outputs (golden) and every gradient of the oracle; frozen tensors stay bit-identical."""
    from ofb_b200.engine import SearchStepEngine
    from ofb_b200.mixup import MixParams
    from step_compare import BF16_TOL, GRAD_MAX_TOL, LOSS_TOL, elementwise_bound_applies, rel, rel_l2
    g, cfg, c, P0, inp = _case(path)
    B, depth, phase = c["B"], cfg.depth, c["phase"]
    eng0 = SearchStepEngine(cfg.embed_dim, cfg.num_heads, depth, B, drop_path_rate=c["dpr"], lr=c["lr"])
    eng0.load_params(P0)
    eng0.set_schedule(c["compress_at"])
    raw = (inp.raw_images if phase == 2 else inp.images).cuda()
    lab = inp.labels.cuda()
    drop_u = ((inp.drop_scale > 0).float().reshape(depth * 2, B) * 0.999).cuda()
    eng0.step(raw, lab, drop_u=drop_u, update=False)           # ranks for the plan
    eng0.grads.zero_()
    eng, finished, executed = eng0.prune_event(0.2)
    assert finished and executed and eng is not eng0 and eng.finish_search and eng.prenorm
    assert (eng.Dv, eng.heads, eng.hdims, eng.hids) == (int(g["embed"]), g["heads"].tolist(), g["head_dims"].tolist(),
                                                        g["hiddens"].tolist())
    eng.set_schedule(c["epoch_frac"])
    if phase == 2:
        eng.enter_post_search()
        eng.set_schedule(c["epoch_frac"])                     # the schedule hook keeps running; it must not undo the reset
        assert eng.keep_ratio == 1.0
        soft = inp.soft_target.cuda()
        # the unmixed batch + MixParams: blend fused into the im2col
        scal = eng.step(raw, None, drop_u=drop_u, update=False, target=soft, mix=MixParams(inp.lam, inp.box))
    else:
        assert abs(eng.keep_ratio - inp.keep_ratio) < 1e-12
        scal = eng.step(raw, lab, noise=inp.noise.cuda(), drop_u=drop_u, update=False)
    torch.cuda.synchronize()
    scal = scal.cpu()
    assert eng.padding_is_clean()
    # --- the reference's own numbers ---
    assert _rel(eng.logits.cpu().numpy(), g["logits"]) < BF16_TOL
    assert _rel(float(scal[0]), g["loss_base"]) < LOSS_TOL and _rel(float(scal[3]), g["loss_total"]) < LOSS_TOL
    assert float(scal[1]) == 0.0                               # no architecture term after finish_search
    if phase == 1:
        assert _rel(float(scal[2]), g["loss_decoder"]) < LOSS_TOL
    else:
        assert float(scal[2]) == 0.0 and float(scal[3]) == float(scal[0])
    # --- every gradient against the oracle on the same parameters ---
    Pp, P1, out, grads, shape, sw, plans = oracle_post(cfg, P0, inp, c)
    got = eng.named_grads()
    worst_l2, worst_max = ("", 0.0), ("", 0.0)
    for k, gr in grads.items():
        if gr is None or float(gr.abs().max()) == 0.0:
            assert float(got[k].abs().max()) == 0.0, k
            continue
        e2, em = rel_l2(got[k], gr), rel(got[k], gr)
        worst_l2 = max(worst_l2, (k, e2), key=lambda kv: kv[1])
        if elementwise_bound_applies(k, B):
            worst_max = max(worst_max, (k, em), key=lambda kv: kv[1])
    print("worst L2", worst_l2, "worst max", worst_max)
    # fixed bound of step_compare.py for EVERY gradient tensor (decoder and scores included)
    assert worst_l2[1] < BF16_TOL and worst_max[1] < GRAD_MAX_TOL
    for key in g.files:                                        # tensors the reference gives no gradient: exact zeros here
        if key.startswith("nograd:") and key[7:] in got:
            assert float(got[key[7:]].abs().max()) == 0.0, key
    # --- update: frozen tensors bit-identical, the rest moves; graph replay of the phase ---
    before = {k: v.detach().clone() for k, v in eng.named_parameters().items()}
    eng.optimizer_step()
    after = eng.named_parameters()
    for key in g.files:
        if key.startswith("nograd:") and key[7:] in before:
            assert torch.equal(before[key[7:]], after[key[7:]]), key
    assert not torch.equal(before["head.weight"], after["head.weight"])
    if phase == 2:
        mixed = inp.images.cuda()
        eng.step_graphed(mixed, None, target=soft)
        eng.step_graphed(mixed, None, target=soft)
    else:
        eng.step_graphed(raw, lab)
    torch.cuda.synchronize()
    assert eng.padding_is_clean() and torch.isfinite(eng.scal).all()
    for k in ("decoder.0.weight", "mask_token") if phase == 2 else ():
        assert torch.equal(before[k], eng.named_parameters()[k]), k
    assert all(torch.equal(before[k], eng.named_parameters()[k]) for k in eng.alpha_names)
