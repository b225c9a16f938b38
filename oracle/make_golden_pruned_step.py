"""TEST INFRASTRUCTURE ONLY — one full search step of the UNMODIFIED reference AFTER a truncating prune event: scripted
alphas make compress() (vision_transformer.py:785-950) slice embedding dims, heads, head channels and hidden units while every
module stays in the search, then the body of engine.search_one_epoch runs on the compressed model (fresh optimizers). Stores
logits, loss terms, FLOPs and gradient fingerprints in tests/golden/pruned_step/<case>.npz and checks the generalised oracle
(ofb_oracle.forward_step(shape=...)) fed with the planner's gathers against them.   Run in the build container only.
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
from fixtures import make_inputs, make_params, pruned_shape_from_plans, search_modules, summarize  # noqa: E402
from make_golden import run_reference  # noqa: E402
from ofb_oracle import ModelCfg, _desc_rank, default_switches, train_step, w_p_schedule  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden", "pruned_step")
CASES = {
    "tiny_d2": dict(D=192, H=3, depth=2, B=8, epoch_frac=6.0, dpr=0.1, lr=1e-3, offset=0),
    "small_d3": dict(D=384, H=6, depth=3, B=8, epoch_frac=11.0, dpr=0.1, lr=1e-3, offset=1),
    "small_d2": dict(D=384, H=6, depth=2, B=8, epoch_frac=2.0, dpr=0.0, lr=1e-3, offset=2),
    # mixed events incl. finalisation (make_golden_prune.script_alphas): finished modules gate with their frozen score, and a
    # finished embedding search switches the blocks to standard pre-norm
    "mixed_tiny_d2": dict(D=192, H=3, depth=2, B=8, epoch_frac=7.0, dpr=0.1, lr=1e-3, offset=0, mixed=True),
    "mixed_small_d3": dict(D=384, H=6, depth=3, B=8, epoch_frac=12.0, dpr=0.1, lr=1e-3, offset=0, mixed=True),
    "mixed_tiny_d3": dict(D=192, H=3, depth=3, B=8, epoch_frac=20.0, dpr=0.0, lr=1e-3, offset=3, mixed=True),
}


def script(P, c):
    if c.get("mixed"):
        from make_golden_prune import script_alphas
        cfg = ModelCfg(embed_dim=c["D"], num_heads=c["H"], depth=c["depth"])
        return script_alphas(P, cfg, offset=c["offset"])
    return script_truncations(P, offset=c["offset"])


def script_truncations(P, offset=0, seed=4):
    """Alphas that make every module either keep everything, lose an interior cell, or lose its trailing rows / columns
    (physical truncation) - never a single survivor. Scores without ties (see make_golden_prune.py)."""
    g = torch.Generator().manual_seed(seed)
    for k in sorted(k for k in P if k.endswith(".score")):
        P[k] = torch.randn(P[k].shape, generator=g) * 0.2
    for n, k in enumerate(sorted(k for k in P if k.endswith(".alpha"))):
        a = torch.rand(P[k].shape, generator=g)
        mode = (n + offset) % 4
        if mode == 0:
            a = a * 0.05
        elif mode == 1:
            a.view(-1)[a.numel() // 2] = -9.0
        elif mode == 2:
            a[..., -2:] = -9.0
        else:
            if a.shape[0] > 1:
                a[-1, :] = -9.0
            a[..., -3:] = -9.0
        P[k] = a
    return P


def plan_on_cpu(cfg, P, switches, thresh=0.2):
    import ofb_b200  # noqa: F401
    from ofb_b200 import prune
    plans, dims = {}, {}
    for prefix, kind, H, dim, widths, heads in search_modules(cfg):
        score = P[prefix + ".score"].reshape(H, dim)
        hr = _desc_rank(torch.sigmoid(score).sum(-1)) if H > 1 else torch.zeros(1, dtype=torch.long)
        plans[prefix] = prune.plan_module(prefix, kind, P[prefix + ".alpha"], switches[prefix], widths, heads, hr,
                                          _desc_rank(score), thresh)
        dims[prefix] = dict(heads=H, dim=dim)
    return plans, dims


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    import ofb_b200  # noqa: F401
    from ofb_b200 import prune
    for name, c in CASES.items():
        cfg = ModelCfg(embed_dim=c["D"], num_heads=c["H"], depth=c["depth"])
        P0 = script(make_params(cfg, seed=0), c)
        inp = make_inputs(cfg, c["B"], seed=1, epoch_frac=c["epoch_frac"], drop_path_rate=c["dpr"])
        sw0 = default_switches(cfg)

        def compress(model):
            for m in model.searchable_modules:
                m.update_w(c["epoch_frac"], 20)
            with contextlib.redirect_stdout(io.StringIO()):
                fin, ex, _, _, _ = model.compress(0.2, None, None, None)
            assert ex and not fin and (c.get("mixed") or not any(m.finish_search for m in model.searchable_modules))

        class AnyShape(dict):            # run_reference copies P0 by name before the hook: shapes still match there
            pass
        with contextlib.redirect_stdout(io.StringIO()):
            ref = run_reference(cfg, P0, inp, sw0, c["dpr"], c["lr"], c["epoch_frac"], after_load=compress)

        # oracle side: planner + gathers (parity with compress() is pinned by make_golden_prune.py), then the generalised step
        plans, dims = plan_on_cpu(cfg, P0, sw0)
        w_p = w_p_schedule(c["epoch_frac"])
        Pp = prune.gather_pruned(plans, {k: v for k, v in P0.items() if k != "alpha_patch"}, dims, w_p)
        Pp["alpha_patch"] = P0["alpha_patch"]
        shape = pruned_shape_from_plans(cfg, plans)
        sw = {k: pl.switch for k, pl in plans.items()}
        out, grads = train_step({k: v.clone() for k, v in Pp.items()}, {}, inp, cfg, lr=c["lr"], step=1, switches=sw, shape=shape)
        checks = {"logits": rel(out.logits, ref["logits"]), "base": rel(out.loss_base, ref["base"]),
                  "arch": rel(out.loss_arch, ref["arch"]), "dec": rel(out.loss_decoder, ref["dec"]),
                  "total": rel(out.loss_total, ref["total"]),
                  "flops_s": abs(float(out.loss_terms["flops_searched"]) - ref["flops"][1]) / ref["flops"][1],
                  "flops_o": abs(float(out.loss_terms["flops_ori"]) - ref["flops"][0]) / ref["flops"][0]}
        for k, g in ref["grads"].items():
            if g is not None:
                checks["grad:" + k] = rel(grads[k], g)
        bad = {k: v for k, v in checks.items() if v > 1e-4}
        print(f"[{name}] shape embed {shape.embed} heads {shape.heads} dims {shape.head_dims} hid {shape.hiddens}; oracle vs "
              f"reference worst rel err {max(checks.values()):.3e} over {len(checks)} tensors; arch {float(ref['arch']):.5f} "
              f"flops {ref['flops']}")
        assert not bad, bad
        gold = {"logits": ref["logits"].numpy(), "loss_base": ref["base"].numpy(), "loss_arch": ref["arch"].numpy(),
                "loss_decoder": ref["dec"].numpy(), "loss_total": ref["total"].numpy(), "flops": np.array(ref["flops"]),
                "cfg": np.array([c["D"], c["H"], c["depth"], c["B"]]), "epoch_frac": np.array(c["epoch_frac"]),
                "dpr": np.array(c["dpr"]), "lr": np.array(c["lr"]), "offset": np.array(c["offset"]),
                "mixed": np.array(bool(c.get("mixed"))),
                "embed": np.array(shape.embed), "heads": np.array(shape.heads), "head_dims": np.array(shape.head_dims),
                "hiddens": np.array(shape.hiddens)}
        for k, g in ref["grads"].items():
            if g is not None:
                gold["gsum:" + k] = summarize(g).numpy()
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **gold)


if __name__ == "__main__":
    main()
