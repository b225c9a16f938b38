"""TEST INFRASTRUCTURE ONLY — runs the UNMODIFIED reference's compress() (vision_transformer.py:785-950 and the module
compress methods) on seeded parameters with scripted architecture parameters (alphas) that exercise every outcome of a prune
event - nothing, switch-only, truncation of channels / heads / hidden units / embedding dims, finalisation - and stores what
the reference leaves behind in tests/golden/prune/<case>.npz: switch cells, alphas, scores and a fingerprint of every tensor.
Run in the build container only:   python oracle/make_golden_prune.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from fixtures import make_params, summarize  # noqa: E402
from ofb_oracle import ModelCfg  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden", "prune")
CASES = {
    "tiny_d2": dict(D=192, H=3, depth=2, epoch_frac=7.0, offset=0),
    "small_d3": dict(D=384, H=6, depth=3, epoch_frac=12.0, offset=0),
    "small_d2_o2": dict(D=384, H=6, depth=2, epoch_frac=3.0, offset=2),
    "tiny_d3_o3": dict(D=192, H=3, depth=3, epoch_frac=20.0, offset=3),
    "small_d2_o3": dict(D=384, H=6, depth=2, epoch_frac=9.0, offset=3),
}


def script_alphas(P, cfg, seed=3, offset=0):
    """Overwrite the alphas so that the modules hit different branches of compress()."""
    g = torch.Generator().manual_seed(seed)
    # importance scores without ties: the step fixtures clamp them at +-2 sigma, which creates equal values, and the order
    # torch.argsort gives equal scores is unspecified (the engine breaks ties towards the lower index; trained scores are
    # continuous, so ties do not occur in a real search)
    for k in sorted(k for k in P if k.endswith(".score")):
        P[k] = torch.randn(P[k].shape, generator=g) * 0.2
    names = sorted(k for k in P if k.endswith(".alpha"))
    for n, k in enumerate(names):
        a = torch.rand(P[k].shape, generator=g)
        mode = (n + offset) % 5
        if mode == 0:                       # nothing dies (flat alphas)
            a = a * 0.05
        elif mode == 1:                     # one interior cell dies -> switch-only (or truncation if it is the last column)
            a.view(-1)[a.numel() // 2] = -9.0
        elif mode == 2:                     # the trailing columns die -> channel / width truncation
            a[..., -2:] = -9.0
        elif mode == 3:                     # attention: last row and last column die; 1-D: last three columns
            if a.shape[0] > 1:
                a[-1, :] = -9.0
            a[..., -3:] = -9.0
        else:                               # a single survivor -> finalise
            a[:] = -9.0
            a.view(-1)[(a.numel() - 1) // 3] = 2.0
        P[k] = a
    return P


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    ref_shim.install()
    for name, c in CASES.items():
        cfg = ModelCfg(embed_dim=c["D"], num_heads=c["H"], depth=c["depth"])
        P0 = script_alphas(make_params(cfg, seed=0), cfg, offset=c["offset"])
        model = ref_shim.build_reference_model(cfg.embed_dim, cfg.num_heads, cfg.depth, 0.1, cfg.num_classes)
        with torch.no_grad():
            for k, p in model.named_parameters():
                p.copy_(P0[k])
        for m in model.searchable_modules:
            m.update_w(c["epoch_frac"], 20)
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            finish, executed, _, _, _ = model.compress(0.2, None, None, None)
        gold = {"cfg": np.array([c["D"], c["H"], c["depth"]]), "epoch_frac": np.array(c["epoch_frac"]), "offset": np.array(c["offset"]),
                "finish": np.array(bool(finish)), "executed": np.array(bool(executed))}
        mods = {"patch_embed": model.patch_embed}
        for l, blk in enumerate(model.blocks):
            mods[f"blocks.{l}.attn"] = blk.attn
            mods[f"blocks.{l}.mlp"] = blk.mlp
        summary = []
        for k, m in mods.items():
            gold["switch:" + k] = m.switch_cell.numpy()
            gold["state:" + k] = np.array([bool(m.finish_search), bool(m.execute_prune)])
            summary.append(f"{k}: alive {int(m.switch_cell.sum())}/{m.switch_cell.numel()} exec={m.execute_prune} "
                           f"fin={m.finish_search} score{tuple(m.score.shape)}")
        for k, p in model.named_parameters():
            gold["shape:" + k] = np.array(p.shape)
            if k.endswith(".alpha") or k.endswith(".score") or k == "alpha_patch":
                gold["full:" + k] = p.detach().numpy()
            else:
                gold["sum:" + k] = summarize(p.detach()).numpy()
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **gold)
        print(f"[{name}] finish={finish} executed={executed}")
        for s in summary:
            print("   ", s)


if __name__ == "__main__":
    main()
