"""Drop-in boundary: the reference's OWN training loop - engine.search_one_epoch (engine.py:75-219), its
OFBSearchLOSS, its three optim.AdamW optimizers and its compress() - driving a model whose searchable modules were created
through the patched ModuleInjection factory and whose forward / backward run on the sm_100a engine (ofb_b200.modules), against
the same loop on the plain, unmodified reference model. The reference is imported from the staged copy oracle/_ref (or
/root/reference in the build container)."""
import contextlib
import io
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import ref_shim  # noqa: E402

pytestmark = pytest.mark.gpu


class _Sched:
    def step_update(self, n):
        pass


def _build(D, H, depth, dev, injected):
    import models.layers as L
    import models.vision_transformer as VT
    from ofb_b200 import modules
    if injected:
        modules.install(L, VT)
    else:
        modules.uninstall()
    model = ref_shim.build_reference_model(D, H, depth, drop_path_rate=0.1, seed=0).to(dev)
    if injected:
        names = {type(model.patch_embed).__name__, type(model.blocks[0].attn).__name__, type(model.blocks[0].mlp).__name__}
        assert names == {"OFBPatchEmbed", "OFBSparseAttention", "OFBSparseMlp"}, names
    return model


def _script(model):
    """Alphas that make the first compress() slice tensors: embedding and one MLP lose their widest candidates, one attention
    module its widest head-channel column; one interior cell elsewhere only switches off."""
    with torch.no_grad():
        model.patch_embed.alpha[0, -2:] = -9.0
        model.blocks[0].mlp.alpha[0, -1] = -9.0
        model.blocks[1].attn.alpha[:, -1] = -9.0
        model.blocks[0].attn.alpha[0, 2] = -9.0


def _run_epoch(model, batches, dev, accum=1):
    import engine as ref_engine           # the reference's engine.py
    import ref_runner
    ddp = ref_shim.FakeDDP(model)
    opt_p, opt_d, opt_a = ref_runner.build_optimizers(model, lr=1e-3)
    crit = ref_runner.build_criterion(dev)
    args = types.SimpleNamespace(accum_iter=accum, warmup_epochs=20, epochs=100)
    torch.manual_seed(11)
    torch.cuda.manual_seed(11)
    with contextlib.redirect_stdout(io.StringIO()):
        stats, finish, executed, *_ = ref_engine.search_one_epoch(
            ddp, crit, 1.0, batches, opt_p, opt_d, opt_a, _Sched(), _Sched(), _Sched(), dev, epoch=3, args=args)
    torch.cuda.synchronize()
    return stats, finish, executed


@pytest.mark.parametrize("D,H,depth,B", [(192, 3, 2, 16)])
def test_reference_loop_on_injected_modules(cuda_dev, D, H, depth, B):
    if not ref_shim.available():
        pytest.skip("unmodified reference not staged (oracle/make_ref.py)")
    ref_shim.install()
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(5)
    batches = [(torch.randn(B, 3, 224, 224, generator=g), torch.randint(0, 1000, (B,), generator=g)) for _ in range(6)]

    plain = _build(D, H, depth, dev, injected=False)
    _script(plain)
    init = {k: v.detach().clone() for k, v in plain.state_dict().items()}
    fused = _build(D, H, depth, dev, injected=True)
    fused.load_state_dict(init)

    # eval-mode forward (engine.evaluate, engine.py:222-257) at the initial parameters
    plain.eval(); fused.eval()
    x = batches[0][0].to(dev)
    with torch.no_grad():
        lp, _ = plain(x.clone())
        lf, aux = fused(x.clone())
    assert aux == (0., None)
    assert float((lf - lp).abs().max() / lp.abs().max()) < 2e-2

    # 6 steps of the reference's search_one_epoch; len(loader) // 3 = 2 -> compress() after steps 2, 4, 6; the first one slices
    sp, fin_p, ex_p = _run_epoch(plain, batches, dev)
    sf, fin_f, ex_f = _run_epoch(fused, batches, dev)
    print("plain :", {k: round(v, 5) for k, v in sp.items()})
    print("fused :", {k: round(v, 5) for k, v in sf.items()})
    assert (fin_p, ex_p) == (fin_f, ex_f) and ex_p
    for k in ("loss_param", "loss_total", "loss_arch", "loss_decoder"):
        assert abs(sf[k] - sp[k]) <= 5e-3 * abs(sp[k]), (k, sf[k], sp[k])
    # the prune decisions are exact: same shapes, same switch cells, same surviving search space
    pp, pf = dict(plain.named_parameters()), dict(fused.named_parameters())
    assert {k: tuple(v.shape) for k, v in pp.items()} == {k: tuple(v.shape) for k, v in pf.items()}
    assert plain.pos_embed.shape[-1] < D, "the scripted alphas were meant to truncate the embedding"
    for mp, mf in zip(plain.searchable_modules, fused.searchable_modules):
        assert torch.equal(mp.switch_cell.cpu(), mf.switch_cell.cpu())
        assert (mp.finish_search, mp.execute_prune) == (mf.finish_search, mf.execute_prune)
    # parameters after six AdamW updates: Adam normalises the gradient, so an entry whose gradient is within the bf16 error of
    # zero can land a few lr apart - the bulk must agree. Alphas (driven by the fp32 architecture loss) agree closely.
    worst = ("", 1.0)
    for k in pp:
        a, b = pp[k].detach().float().flatten(), pf[k].detach().float().flatten()
        if k.endswith(".alpha"):
            assert float((a - b).abs().max()) < 2e-3, k
        if a.numel() < 256:
            continue
        ia = init[k].to(dev).float()
        if ia.shape != pp[k].shape:
            # sliced by compress(): rows / columns are gathered in RANK order of scores that differ in their last bits between
            # the two runs, so the element order is not comparable - the shapes (checked above) are
            continue
        da, db = a - ia.flatten(), b - ia.flatten()
        if float(da.norm()) == 0:
            continue
        cos = float(torch.nn.functional.cosine_similarity(da, db, dim=0))
        worst = min(worst, (k, cos), key=lambda kv: kv[1])
    print("worst cosine similarity of a parameter update:", worst)
    assert worst[1] > 0.9, worst

    # per-module forward has no eager fallback
    from ofb_b200._lib import OfbError
    with pytest.raises(OfbError):
        fused.blocks[0].mlp(torch.zeros(1, 197, fused.pos_embed.shape[-1], device=dev))
    from ofb_b200 import modules
    modules.uninstall()


def test_injected_step_gradients_match_plain_reference(cuda_dev):
    """One forward / criterion / backward of the reference's loop body on both models with identical draws: logits, losses and
    every parameter gradient (the fixed bound of step_compare.py: rel-L2 2e-2)."""
    if not ref_shim.available():
        pytest.skip("unmodified reference not staged (oracle/make_ref.py)")
    from step_compare import BF16_TOL, LOSS_TOL, rel, rel_l2
    import ref_runner
    ref_shim.install()
    dev = torch.device("cuda")
    D, H, depth, B = 384, 6, 3, 32
    g = torch.Generator().manual_seed(9)
    x, y = torch.randn(B, 3, 224, 224, generator=g).to(dev), torch.randint(0, 1000, (B,), generator=g).to(dev)
    out = {}
    for name, injected in (("plain", False), ("fused", True)):
        model = _build(D, H, depth, dev, injected)
        if name == "plain":
            init = {k: v.detach().clone() for k, v in model.state_dict().items()}
        else:
            model.load_state_dict(init)
        model.train()
        model.adjust_masking_ratio(4.0, 20, 100, max_ratio=0.95, min_ratio=0.75)
        for m in model.searchable_modules:
            m.update_w(4.0, 20)
        ddp = ref_shim.FakeDDP(model)
        crit = ref_runner.build_criterion(dev)
        torch.manual_seed(3)
        torch.cuda.manual_seed(3)
        with contextlib.redirect_stdout(io.StringIO()):
            logits, (dec, _) = ddp(x.clone())
            base, arch = crit(x, logits, y, ddp, "arch", 1.0, False)
        total = base + arch + (base / dec).data.clone() * dec
        total.backward()
        out[name] = dict(logits=logits.detach(), base=base.detach(), arch=arch.detach(), dec=dec.detach(),
                         grads={k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None})
    from ofb_b200 import modules
    modules.uninstall()
    p, f = out["plain"], out["fused"]
    assert rel(f["logits"], p["logits"]) < BF16_TOL
    assert rel(f["base"], p["base"]) < LOSS_TOL and rel(f["dec"], p["dec"]) < LOSS_TOL and rel(f["arch"], p["arch"]) < 1e-4
    assert set(p["grads"]) == set(f["grads"])
    worst = max(((k, rel_l2(f["grads"][k], p["grads"][k])) for k in p["grads"]), key=lambda kv: kv[1])
    print("worst gradient (rel L2):", worst)
    assert worst[1] < BF16_TOL, worst
