"""Eval-mode forward (engine.evaluate path of the reference, engine.py:222-257) — the CPU oracle against the golden vectors
of the unmodified reference (CPU), and the GPU engine's evaluate() against both (gpu marker)."""
import glob
import os

import numpy as np
import pytest
import torch

from fixtures import make_inputs, make_params
from ofb_oracle import ModelCfg, default_switches, forward_step

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval", "*.npz")))
IDS = [os.path.basename(p)[:-4] for p in GOLD]


def _case(path):
    g = np.load(path)
    D, H, depth, B = (int(x) for x in g["cfg"])
    cfg = ModelCfg(embed_dim=D, num_heads=H, depth=depth)
    P = make_params(cfg, seed=0)
    inp = make_inputs(cfg, B, seed=1, epoch_frac=float(g["epoch_frac"]), drop_path_rate=0.0, keep_ratio=1.0)
    sw = default_switches(cfg)
    for k in sw:
        sw[k] = torch.from_numpy(g["switch:" + k])
    return g, cfg, P, inp, sw


def _metrics(logits, labels):
    """CrossEntropyLoss + timm accuracy (top-k hit = fewer than k logits ahead of the label's)."""
    logp = torch.log_softmax(logits.double(), dim=-1)
    nll = -logp.gather(1, labels.unsqueeze(1)).squeeze(1)
    ly = logits.gather(1, labels.unsqueeze(1))
    ahead = (logits > ly).sum(1)
    return nll.mean(), (ahead < 1).double().mean(), (ahead < 5).double().mean()


@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_oracle_eval_matches_reference_golden(path):
    g, cfg, P, inp, sw = _case(path)
    out = forward_step(P, inp, cfg, sw)
    ref = torch.from_numpy(g["logits"])
    assert float((out.logits - ref).abs().max() / ref.abs().max()) < 1e-4
    assert float(out.loss_decoder) == 0.0                     # no PMIM branch in eval mode (vt:719)
    loss, a1, a5 = _metrics(out.logits.detach(), torch.from_numpy(g["labels"]))
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    assert abs(float(a1) * 100 - float(g["acc1"])) < 1e-4 and abs(float(a5) * 100 - float(g["acc5"])) < 1e-4   # fp32 percentages


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_engine_evaluate_matches_reference_golden(cuda_dev, path):
    from ofb_b200.engine import SearchStepEngine
    from step_compare import BF16_TOL, LOSS_TOL
    g, cfg, P, inp, sw = _case(path)
    B = inp.images.shape[0]
    eng = SearchStepEngine(cfg.embed_dim, cfg.num_heads, cfg.depth, B, drop_path_rate=0.1, switches=sw)
    eng.load_params(P)
    eng.set_schedule(float(g["epoch_frac"]))
    labels = torch.from_numpy(g["labels"]).cuda()
    # a training step first: evaluate() must not depend on leftovers of the PMIM mask / DropPath state
    eng.step(inp.images.cuda(), inp.labels.cuda(), update=False)
    eng.grads.zero_()
    out = eng.evaluate(inp.images.cuda(), labels).cpu()
    torch.cuda.synchronize()
    logits = eng.logits.cpu()
    ref = torch.from_numpy(g["logits"])
    assert float((logits - ref).abs().max() / ref.abs().max()) < BF16_TOL
    # the metric kernel is exact on the engine's own logits
    loss, a1, a5 = _metrics(logits, labels.cpu())
    assert abs(float(out[0]) - float(loss)) < 1e-5 * abs(float(loss))
    assert abs(float(out[1]) - float(a1)) < 1e-6 and abs(float(out[2]) - float(a5)) < 1e-6
    # against the reference's numbers: the loss within the bf16 loss tolerance; a hit flag may only differ where the
    # reference's deciding logit margin is below the bf16 logit tolerance
    assert abs(float(out[0]) - float(g["loss"])) < LOSS_TOL * abs(float(g["loss"]))
    lab = labels.cpu()
    ly = ref.gather(1, lab.unsqueeze(1))
    srt = ref.sort(dim=1, descending=True).values
    tol = 2 * BF16_TOL * float(ref.abs().max())
    for k, col in ((1, 1), (5, 2)):
        ref_hit = ((ref > ly).sum(1) < k)
        got_hit = ((logits > logits.gather(1, lab.unsqueeze(1))).sum(1) < k)
        margin = torch.minimum((ly.squeeze(1) - srt[:, k]).abs(), (ly.squeeze(1) - srt[:, k - 1]).abs())
        decided = margin > tol
        assert torch.equal(ref_hit[decided], got_hit[decided]), f"top-{k}"
