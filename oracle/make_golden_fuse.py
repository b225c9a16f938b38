"""TEST INFRASTRUCTURE ONLY — the hand-off from the search to the finetune / fused-inference model on the UNMODIFIED reference:
scripted alphas finalise every searchable module in one compress() (vision_transformer.py:785-950), fuse() folds the frozen
gates into the weights (vt:747-757), and the fused model is run in eval mode. Stores the fused tensors' fingerprints, the
per-layer pruned dims and the eval logits in tests/golden/fuse/<case>.npz.   Run in the build container only.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from fixtures import make_inputs, make_params, summarize  # noqa: E402
from ofb_oracle import ModelCfg  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden", "fuse")
CASES = {"tiny_d2": dict(D=192, H=3, depth=2, B=4, epoch_frac=15.0), "small_d3": dict(D=384, H=6, depth=3, B=3, epoch_frac=20.0)}


def script_single_survivor(P, seed=5):
    """Every alpha table keeps exactly one cell (a different one per module); scores without ties (see make_golden_prune)."""
    g = torch.Generator().manual_seed(seed)
    for k in sorted(k for k in P if k.endswith(".score")):
        P[k] = torch.randn(P[k].shape, generator=g) * 0.2
    for k in sorted(k for k in P if k.endswith(".alpha")):
        a = torch.full(P[k].shape, -9.0)
        a.view(-1)[int(torch.randint(0, a.numel(), (1,), generator=g))] = 2.0
        P[k] = a
    return P


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    ref_shim.install()
    for name, c in CASES.items():
        cfg = ModelCfg(embed_dim=c["D"], num_heads=c["H"], depth=c["depth"])
        P0 = script_single_survivor(make_params(cfg, seed=0))
        model = ref_shim.build_reference_model(cfg.embed_dim, cfg.num_heads, cfg.depth, 0.1, cfg.num_classes)
        with torch.no_grad():
            for k, p in model.named_parameters():
                p.copy_(P0[k])
        for m in model.searchable_modules:
            m.update_w(c["epoch_frac"], 20)
        with contextlib.redirect_stdout(io.StringIO()):
            finish, executed, _, _, _ = model.compress(0.2, None, None, None)
            assert finish and executed
            model.fuse()
        model.eval()
        inp = make_inputs(cfg, c["B"], seed=1, epoch_frac=c["epoch_frac"], drop_path_rate=0.0, keep_ratio=1.0)
        with torch.no_grad():
            logits, _ = model(inp.images.clone())
        gold = {"cfg": np.array([c["D"], c["H"], c["depth"], c["B"]]), "epoch_frac": np.array(c["epoch_frac"]),
                "logits": logits.numpy(),
                "embed": np.array(model.cls_token.shape[-1]),
                "heads": np.array([b.attn.head_num for b in model.blocks]),
                "head_dims": np.array([b.attn.qkv.out_features // (3 * b.attn.head_num) for b in model.blocks]),
                "hiddens": np.array([b.mlp.fc1.out_features for b in model.blocks])}
        for k, p in model.named_parameters():
            if k.endswith(".score") or k.endswith(".alpha") or k in ("alpha_patch", "mask_token") or k.startswith("decoder."):
                continue
            gold["shape:" + k] = np.array(p.shape)
            gold["sum:" + k] = summarize(p.detach()).numpy()
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **gold)
        print(f"[{name}] embed {int(gold['embed'])} heads {gold['heads'].tolist()} head_dims {gold['head_dims'].tolist()} "
              f"hiddens {gold['hiddens'].tolist()} logits l2 {float(logits.norm()):.4f}")


if __name__ == "__main__":
    main()
