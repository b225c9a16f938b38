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
from make_golden_pruned_step import plan_on_cpu, script_truncations
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
    P0 = script_truncations(make_params(cfg, seed=0), offset=int(g["offset"]))
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
