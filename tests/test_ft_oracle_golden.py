"""The finetune-step CPU oracle (oracle/ft_oracle.py) against the golden vectors produced by the unmodified reference
classes (oracle/make_golden_ft.py: plain VisionTransformer on pruned shapes + lr_decay groups + torch AdamW). Runs without
/root/reference and without a GPU."""
import glob
import os

import numpy as np
import pytest
import torch

from fixtures import summarize
from ft_oracle import SubnetCfg, ft_group, ft_param_shapes, ft_train_step, layer_id, make_ft_inputs, make_ft_params

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ft_*.npz")))


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def load_case(path):
    g = np.load(path)
    cfg = SubnetCfg(embed_dim=int(g["D"]), heads=[int(x) for x in g["heads"]], head_dims=[int(x) for x in g["head_dims"]],
                    hiddens=[int(x) for x in g["hiddens"]])
    P = make_ft_params(cfg, seed=0)
    train = bool(g["train"])
    images, labels, drop_scale, target = make_ft_inputs(cfg, int(g["B"]), seed=1,
                                                        drop_path_rate=float(g["dpr"]) if train else 0.0, soft=bool(g["soft"]))
    return g, cfg, P, images, labels, drop_scale, target


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_ft_oracle_matches_reference_golden(path):
    g, cfg, P, images, labels, drop_scale, target = load_case(path)
    logits, loss, grads = ft_train_step(P, {}, images, labels, cfg, lr=float(g["lr"]), step=1, drop_scale=drop_scale,
                                        target=target)
    tol = 1e-4   # fp32 rel (north_star)
    assert _rel(logits.numpy(), g["logits"]) < tol
    assert _rel(loss.numpy(), g["loss"]) < tol
    n = 0
    for key in g.files:
        if key.startswith("gsum:"):
            assert _rel(summarize(grads[key[5:]]).numpy(), g[key]) < tol, key
            n += 1
    assert n == len(ft_param_shapes(cfg))
    # post-AdamW parameters: the l2 / sum fingerprints move by ~lr per element, compare at fp32 tolerance
    for key in g.files:
        if key.startswith("psum:"):
            assert _rel(summarize(P[key[5:]]).numpy()[:3], g[key][:3]) < 1e-4, key


def test_layer_decay_groups():
    # lr_decay.py:62-75 / 15-59: embeddings are layer 0, block l is layer l + 1, norm / head the last layer
    assert layer_id("cls_token", 12) == 0 and layer_id("patch_embed.proj.weight", 12) == 0
    assert layer_id("blocks.0.attn.qkv.weight", 12) == 1 and layer_id("blocks.11.mlp.fc2.bias", 12) == 12
    assert layer_id("norm.weight", 12) == 13 and layer_id("head.bias", 12) == 13
    key, sc, wd = ft_group("blocks.3.mlp.fc1.weight", (8, 8), 12, 0.05, 0.95)
    assert key == (4, 1) and abs(sc - 0.95 ** 9) < 1e-12 and wd == 0.05
    assert ft_group("pos_embed", (1, 197, 8), 12, 0.05, 0.95)[2] == 0.0
    assert ft_group("head.bias", (1000,), 12, 0.05, 0.95)[1] == 1.0
