"""Finetune-step parity on the GPU: FinetuneStepEngine (CUDA kernels through the C ABI, pruned per-layer
shapes in the zero-padded layout) against the CPU oracle and against the golden vectors of the unmodified reference."""
import glob
import os

import numpy as np
import pytest
import torch

from ft_compare import compare_ft_step
from step_compare import BF16_TOL, LOSS_TOL

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ft_*.npz")))


@pytest.mark.parametrize("D,heads,dims,hid,B,train,dpr,soft", [
    (252, [4, 6, 2], [40, 64, 56], [768, 1152, 384], 2, False, 0.0, False),        # padded embedding (252 -> 256), padded heads
    (336, [6, 4], [48, 32], [960, 1536], 3, True, 0.1, True),                        # DropPath + Mixup soft targets
    (384, [6, 6], [64, 64], [1536, 1536], 2, False, 0.0, False),                     # unpruned DeiT-S blocks
    (168, [2] * 4, [64, 24, 16, 40], [480, 192, 768, 288], 4, False, 0.0, False),  # DeiT-Tiny-like, very narrow heads
])
def test_ft_step_matches_oracle(cuda_dev, D, heads, dims, hid, B, train, dpr, soft):
    res = compare_ft_step(D, heads, dims, hid, B, train=train, dpr=dpr, soft=soft, verbose=True)
    print(res["summary"])
    assert res["ok"], res["summary"]


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_ft_step_matches_reference_golden(cuda_dev, path):
    from fixtures import summarize
    from ft_oracle import SubnetCfg, make_ft_inputs, make_ft_params
    from ofb_b200.finetune_engine import FinetuneStepEngine
    g = np.load(path)
    cfg = SubnetCfg(embed_dim=int(g["D"]), heads=[int(x) for x in g["heads"]], head_dims=[int(x) for x in g["head_dims"]],
                    hiddens=[int(x) for x in g["hiddens"]])
    B, train, dpr, soft = int(g["B"]), bool(g["train"]), float(g["dpr"]), bool(g["soft"])
    P = make_ft_params(cfg, seed=0)
    images, labels, drop_scale, target = make_ft_inputs(cfg, B, seed=1, drop_path_rate=dpr if train else 0.0, soft=soft)
    eng = FinetuneStepEngine(cfg.embed_dim, cfg.heads, cfg.head_dims, cfg.hiddens, B, lr=float(g["lr"]), training_mode=train,
                             drop_path_rate=dpr)
    eng.load_params(P)
    drop_u = (drop_scale > 0).float().reshape(cfg.depth * 2, B) * 0.999
    scal = eng.step(images.cuda(), labels.cuda(), target.cuda() if target is not None else None, drop_u=drop_u.cuda(),
                    update=False)
    torch.cuda.synchronize()

    def rel(a, b):
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))

    assert rel(eng.logits.cpu().numpy(), g["logits"]) < BF16_TOL
    assert rel(float(scal[0]), g["loss"]) < LOSS_TOL
    grads = eng.named_grads()
    worst = ("", 0.0)
    for key in g.files:
        if key.startswith("gsum:"):
            got = summarize(grads[key[5:]].cpu()).numpy()
            assert abs(got[2] - g[key][2]) / (g[key][2] + 1e-30) < BF16_TOL, key          # l2 norm
            e = float(np.abs(got[3:] - g[key][3:]).max() / (np.abs(g[key][3:]).max() + 1e-30))
            worst = max(worst, (key, e), key=lambda kv: kv[1])
    print("worst sampled gradient error:", worst)
    assert worst[1] < (2.5 if cfg.depth >= 12 else 2.0) * BF16_TOL
