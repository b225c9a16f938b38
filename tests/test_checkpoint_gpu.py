"""Checkpoint / resume in the state-dict format with shape metadata (the reference pickles the whole model object,
search.py:671-740, 302-372): an engine rebuilt from state_dict() continues exactly where the original does - unpruned, after a
truncating prune event (pruned shapes + remaining search space in the metadata) and in the post-search phase."""
import io

import pytest
import torch

from fixtures import make_inputs, make_params
from make_golden_fuse import script_single_survivor
from make_golden_pruned_step import script
from ofb_oracle import ModelCfg

pytestmark = pytest.mark.gpu


def _roundtrip(sd):
    buf = io.BytesIO()
    torch.save(sd, buf)
    buf.seek(0)
    return torch.load(buf, weights_only=False)


def _same(a, b, tol_abs, what):
    # every gradient reduction of the step runs in a fixed order (deterministic split-K weight gradients, per-warp column sums of
    # the attention backward combined in order, ticketed bf16 column sums): a restored engine continues BIT FOR BIT
    worst = max(((float((a.p(k).float() - b.p(k).float()).abs().max()), k) for k in a.offsets))
    assert worst[0] <= tol_abs, (what, worst)


@pytest.mark.parametrize("mode", ["unpruned", "pruned", "post"])
def test_resume_continues_identically(cuda_dev, mode):
    from ofb_b200.engine import SearchStepEngine
    cfg = ModelCfg(embed_dim=192, num_heads=3, depth=2)
    B = 2
    P0 = make_params(cfg, seed=0)
    if mode == "pruned":
        P0 = script(P0, dict(D=192, H=3, depth=2, offset=0, mixed=False))
    elif mode == "post":
        P0 = script_single_survivor(P0)
    inp = make_inputs(cfg, B, seed=1, epoch_frac=6.0, drop_path_rate=0.0)
    img, lab, noise = inp.images.cuda(), inp.labels.cuda(), inp.noise.cuda()
    eng = SearchStepEngine(192, 3, 2, B, drop_path_rate=0.0, lr=1e-3)
    eng.load_params(P0)
    eng.set_schedule(6.0)
    soft = None
    for _ in range(2):
        eng.step(img, lab, noise=noise)
    if mode != "unpruned":
        eng, finished, executed = eng.prune_event(0.2)
        assert executed and finished == (mode == "post")
        if mode == "post":
            eng.enter_post_search()
            soft = torch.softmax(torch.randn(B, 1000, device="cuda"), -1)
        eng.step(img, lab, noise=noise, target=soft)
    sd = _roundtrip(eng.state_dict())
    assert sd["format"] == "ofb_b200.search/1" and all(not v.is_cuda for v in sd["params"].values())
    if mode != "unpruned":
        assert sd["pruned"] is not None and sd["pruned"]["embed"] == eng.Dv
    res = SearchStepEngine.from_state_dict(sd, batch=B)
    assert (res.Dv, res.heads, res.hdims, res.hids, res.step_count) == (eng.Dv, eng.heads, eng.hdims, eng.hids, eng.step_count)
    assert res.finish_search == eng.finish_search and res.decoder_frozen == eng.decoder_frozen and res.keep_ratio == eng.keep_ratio
    _same(res, eng, 0.0, "restore")                            # the restored parameters are the saved ones, bit for bit
    assert torch.equal(res.adam_v, eng.adam_v) and torch.equal(res.adam_m, eng.adam_m) and res.padding_is_clean()
    for e in (eng, res):
        e.step(img, lab, noise=noise, target=soft)
        e.step_graphed(img, lab, target=soft) if mode == "post" else e.step(img, lab, noise=noise)
    torch.cuda.synchronize()
    if mode == "post":
        # the post-search leg replays a CUDA graph, whose PMIM / DropPath draws come from torch's generator (different offsets in
        # the two engines' histories do not matter here: PMIM is off and the DropPath rate is 0)
        pass
    _same(res, eng, 0.0, "after two more steps")
    assert torch.equal(res.scal[:4], eng.scal[:4])


def test_rejects_foreign_checkpoint(cuda_dev):
    from ofb_b200.engine import SearchStepEngine
    with pytest.raises(ValueError):
        SearchStepEngine.from_state_dict({"format": "something else"}, batch=2)
