"""TEST INFRASTRUCTURE ONLY — the POST-SEARCH phase of engine.search_one_epoch on the UNMODIFIED reference.
Scripted alphas make compress() finalise every searchable module (finish_search=True, engine.py:203-208: the architecture
optimizer is dropped, the criterion returns the base loss alone, losses.py:105-106). Two sub-phases are run:
  p1  the rest of that epoch: PMIM masking + decoder still on, label-smoothing CE, optimizer_param + optimizer_decoder;
  p2  from the next epoch on (search.py:641-656): reset_mask_ratio(1.0) (no masking -> no decoder branch, vt:595-612, 719),
      freeze_decoder() (mask_token + decoder frozen, vt:534-539), Mixup / CutMix soft targets with timm SoftTargetCrossEntropy,
      optimizer_param alone.
The mixed batch itself comes from oracle/ofb_oracle.mixup_batch (timm.data.Mixup is third-party and not installed: see the
note there); the reference model, criterion and optimizers run unmodified on it. Stores logits, losses, gradient and
post-update parameter fingerprints in tests/golden/post/<case>.npz and checks the oracle (train_step(finish_search=True))
against them.   Run in the build container only:   python oracle/make_golden_post.py
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_shim  # noqa: E402
from fixtures import make_inputs, make_params, pruned_shape_from_plans, summarize  # noqa: E402
from make_golden_fuse import script_single_survivor  # noqa: E402
from make_golden_pruned_step import plan_on_cpu  # noqa: E402
from ofb_oracle import (ModelCfg, adamw_step, default_switches, group_hparams, mixup_batch, param_group, train_step,  # noqa: E402
                        w_p_schedule)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden", "post")
CASES = {
    # compress_at: epoch fraction of the finalising prune event (fixes w_p inside the finalised scores)
    "p1_tiny_d2": dict(D=192, H=3, depth=2, B=8, phase=1, compress_at=15.0, epoch_frac=15.4, dpr=0.1, lr=1e-3),
    "p2_tiny_d2_mix": dict(D=192, H=3, depth=2, B=8, phase=2, compress_at=15.0, epoch_frac=25.0, dpr=0.1, lr=1e-3, lam=0.37,
                           box=None),
    "p2_small_d3_cut": dict(D=384, H=6, depth=3, B=8, phase=2, compress_at=20.0, epoch_frac=21.0, dpr=0.0, lr=5e-4,
                            lam=None, box=(40, 152, 96, 224)),
}
FROZEN_P2 = ("alpha_patch", "mask_token", "decoder.0.weight", "decoder.0.bias")


def case_inputs(cfg, c):
    """Seeded inputs of a case; phase 2 mixes the batch (box given: CutMix with lam corrected to the box area)."""
    keep = None if c["phase"] == 1 else 1.0
    inp = make_inputs(cfg, c["B"], seed=1, epoch_frac=c["epoch_frac"], drop_path_rate=c["dpr"], keep_ratio=keep)
    if c["phase"] == 2:
        box = c["box"]
        lam = c["lam"] if box is None else 1. - (box[1] - box[0]) * (box[3] - box[2]) / float(cfg.img * cfg.img)
        inp.raw_images, inp.lam, inp.box = inp.images, lam, box
        inp.images, inp.soft_target = mixup_batch(inp.images, inp.labels, lam, box, cfg.num_classes, 0.1)
    return inp


def run_reference_post(cfg, P0, inp, c):
    ref_shim.install()
    import optim as ref_optim
    from losses import DistillationLoss, OFBSearchLOSS
    model = ref_shim.build_reference_model(cfg.embed_dim, cfg.num_heads, cfg.depth, c["dpr"], cfg.num_classes)
    with torch.no_grad():
        for k, p in model.named_parameters():
            p.copy_(P0[k])
    for m in model.searchable_modules:
        m.update_w(c["compress_at"], 20)
    with contextlib.redirect_stdout(io.StringIO()):
        finish, executed, _, _, _ = model.compress(0.2, None, None, None)
    assert finish and executed
    base_criterion = ref_shim.LabelSmoothingCrossEntropy(0.1)
    if c["phase"] == 2:                       # search.py:641-656
        model.reset_mask_ratio(1.0)
        model.freeze_decoder()
        base_criterion = ref_shim.SoftTargetCrossEntropy()
    model.train()
    ddp = ref_shim.FakeDDP(model)
    groups = {"param_nd": [], "param_d": [], "dec_nd": [], "dec_d": [], "arch": []}
    names = {k: [] for k in groups}
    skip = model.no_weight_decay()
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        if len(p.shape) == 1 or name.endswith(".bias") or any(e in name for e in skip):
            key = "param_nd" if "decoder" not in name else "dec_nd"
        elif "alpha" in name:
            key = "arch"
        else:
            key = "param_d" if "decoder" not in name else "dec_d"
        groups[key].append(p)
        names[key].append(name)
    lr = c["lr"]
    opt_param = ref_optim.AdamW([{"params": groups["param_nd"], "weight_decay": 0.},
                                 {"params": groups["param_d"], "weight_decay": 1e-3}],
                                {0: names["param_nd"], 1: names["param_d"]}, lr=lr, eps=1e-8, betas=(0.9, 0.999))
    opt_dec = None
    if c["phase"] == 1:
        opt_dec = ref_optim.AdamW([{"params": groups["dec_nd"], "weight_decay": 0.},
                                   {"params": groups["dec_d"], "weight_decay": 1e-3}],
                                  {0: names["dec_nd"], 1: names["dec_d"]}, lr=lr, eps=1e-8, betas=(0.9, 0.999))
    else:
        assert not groups["dec_nd"] and not groups["dec_d"] and "mask_token" not in names["param_nd"]
    criterion = OFBSearchLOSS(DistillationLoss(base_criterion, None, "none", 0.5, 1.0), torch.device("cpu"),
                              attn_w=0.5, mlp_w=0.5, patch_w=0, embedding_w=0.5, flops_w=5, entropy=True, var=True, norm=True)
    # engine.py:100-117: the schedule hooks still run every step
    model.adjust_masking_ratio(c["epoch_frac"], 20, 100, max_ratio=0.95, min_ratio=0.75)
    for m in model.searchable_modules:
        assert m.finish_search
    queue = ([inp.noise] if c["phase"] == 1 else []) + [u.reshape(-1, 1, 1) for u in inp.drop_draws]
    real_rand = torch.rand

    def fake_rand(*a, **k):
        t = queue.pop(0)
        shape = tuple(a[0]) if len(a) == 1 and isinstance(a[0], (tuple, list, torch.Size)) else tuple(a)
        assert tuple(t.shape) == shape, (t.shape, shape)
        return t.clone()

    torch.rand = fake_rand
    try:
        outputs, (dec_loss, score_loss) = ddp(inp.images.clone())
    finally:
        torch.rand = real_rand
    assert not queue and score_loss is None
    targets = inp.soft_target if c["phase"] == 2 else inp.labels
    loss = criterion(inp.images, outputs, targets, ddp, "arch", cfg.target_flops, True)      # finish_search=True
    assert not isinstance(loss, tuple)
    base_loss, total = loss.item(), loss
    if isinstance(dec_loss, float):
        assert c["phase"] == 2 and dec_loss == 0.
        dec_loss = torch.zeros(())
    else:
        w_dec = (base_loss / dec_loss).data.clone()                                         # engine.py:140-142
        total = total + w_dec * dec_loss
    total.backward()
    grads = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in model.named_parameters()}
    opt_param.step()
    if opt_dec is not None:
        opt_dec.step()
    new_params = {k: p.detach().clone() for k, p in model.named_parameters()}
    trainable = {k for k, p in model.named_parameters() if p.requires_grad}
    return dict(logits=outputs.detach(), base=loss.detach(), dec=dec_loss.detach(), total=total.detach(), grads=grads,
                new_params=new_params, trainable=trainable)


def oracle_post(cfg, P0, inp, c):
    """Oracle side of a case: CPU planner + gathers of the finalising event, then the post-search step."""
    import ofb_b200  # noqa: F401
    from ofb_b200 import prune
    sw0 = default_switches(cfg)
    plans, dims = plan_on_cpu(cfg, P0, sw0)
    assert all(pl.finished for pl in plans.values())
    Pp = prune.gather_pruned(plans, {k: v for k, v in P0.items() if k != "alpha_patch"}, dims, w_p_schedule(c["compress_at"]))
    Pp["alpha_patch"] = P0["alpha_patch"]
    shape = pruned_shape_from_plans(cfg, plans)
    sw = {k: pl.switch for k, pl in plans.items()}
    frozen = FROZEN_P2 if c["phase"] == 2 else ("alpha_patch",)
    P1 = {k: v.clone() for k, v in Pp.items()}
    out, grads = train_step(P1, {}, inp, cfg, lr=c["lr"], step=1, switches=sw, shape=shape, frozen=frozen, finish_search=True)
    return Pp, P1, out, grads, shape, sw, plans


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, c in CASES.items():
        cfg = ModelCfg(embed_dim=c["D"], num_heads=c["H"], depth=c["depth"])
        P0 = script_single_survivor(make_params(cfg, seed=0))
        inp = case_inputs(cfg, c)
        with contextlib.redirect_stdout(io.StringIO()):
            ref = run_reference_post(cfg, P0, inp, c)
        Pp, P1, out, grads, shape, sw, plans = oracle_post(cfg, P0, inp, c)
        checks = {"logits": rel(out.logits, ref["logits"]), "base": rel(out.loss_base, ref["base"]),
                  "total": rel(out.loss_total, ref["total"])}
        if c["phase"] == 1:
            checks["dec"] = rel(out.loss_decoder, ref["dec"])
        n_grad = 0
        for k, g in ref["grads"].items():
            if g is None:
                # no gradient on the reference side -> none (or an exact zero) on ours, and no update
                assert grads.get(k) is None or float(grads[k].abs().max()) == 0.0, k
                continue
            checks["grad:" + k] = rel(grads[k], g)
            n_grad += 1
            pk, mk, vk = Pp[k].clone(), torch.zeros_like(Pp[k]), torch.zeros_like(Pp[k])
            adamw_step(pk, g, mk, vk, 1, **group_hparams(param_group(k, pk), c["lr"]))
            checks["new:" + k] = rel(pk, ref["new_params"][k])
        for k in ref["new_params"]:
            if ref["grads"][k] is None or k not in ref["trainable"]:
                assert torch.equal(ref["new_params"][k], Pp[k].reshape(ref["new_params"][k].shape)), k   # untouched
        bad = {k: v for k, v in checks.items() if v > (1e-6 if k.startswith("new:") else 1e-4)}
        print(f"[{name}] embed {shape.embed} heads {shape.heads} dims {shape.head_dims} hid {shape.hiddens}; oracle vs reference "
              f"worst rel err {max(checks.values()):.3e} over {len(checks)} tensors ({n_grad} gradients); base "
              f"{float(ref['base']):.5f} dec {float(ref['dec']):.5f} total {float(ref['total']):.5f}")
        assert not bad, bad
        gold = {"logits": ref["logits"].numpy(), "loss_base": ref["base"].numpy(), "loss_decoder": ref["dec"].numpy(),
                "loss_total": ref["total"].numpy(), "cfg": np.array([c["D"], c["H"], c["depth"], c["B"]]),
                "phase": np.array(c["phase"]), "compress_at": np.array(c["compress_at"]), "epoch_frac": np.array(c["epoch_frac"]),
                "dpr": np.array(c["dpr"]), "lr": np.array(c["lr"]),
                "lam": np.array(float(getattr(inp, "lam", 1.0))), "box": np.array(getattr(inp, "box", None) or (0, 0, 0, 0)),
                "embed": np.array(shape.embed), "heads": np.array(shape.heads), "head_dims": np.array(shape.head_dims),
                "hiddens": np.array(shape.hiddens)}
        for k, g in ref["grads"].items():
            if g is not None:
                gold["gsum:" + k] = summarize(g).numpy()
                gold["psum:" + k] = summarize(ref["new_params"][k]).numpy()
            else:
                gold["nograd:" + k] = np.array(1)
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **gold)


if __name__ == "__main__":
    main()
