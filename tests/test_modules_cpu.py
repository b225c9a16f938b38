"""Module-surface drop-in, host side (no GPU): the patched ModuleInjection factory builds the OFB* subclasses of the
reference's searchable modules, they keep the reference's attribute / method contract, their weighted_mask (what
OFBSearchLOSS -> get_flops() reads, with its graph to alpha) equals what the reference's own forward computes, the model
pickles, and nothing falls back to the CPU."""
import io
import os
import pickle
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import ref_shim  # noqa: E402

needs_ref = pytest.mark.skipif(not ref_shim.available(), reason="unmodified reference not staged (oracle/make_ref.py)")


@pytest.fixture()
def injected():
    ref_shim.install()
    import models.layers as L
    import models.vision_transformer as VT
    import ofb_b200  # noqa: F401
    from ofb_b200 import modules
    modules.install(L, VT)
    yield L, VT, modules
    modules.uninstall()


@needs_ref
def test_factory_builds_subclasses_with_reference_contract(injected):
    L, VT, modules = injected
    model = ref_shim.build_reference_model(192, 3, depth=2)
    assert isinstance(model.patch_embed, L.MAEPatchEmbed) and type(model.patch_embed).__name__ == "OFBPatchEmbed"
    assert isinstance(model.blocks[0].attn, L.MAESparseAttention) and isinstance(model.blocks[1].mlp, L.MAESparseMlp)
    assert len(model.searchable_modules) == 5 and len(L.ModuleInjection.searchable_modules) == 5
    for m in model.searchable_modules:
        for attr in ("alpha", "score", "switch_cell", "mask", "w_p", "finish_search", "execute_prune", "fused", "update_w",
                     "compress", "fuse", "get_alpha", "get_weight", "get_flops", "get_params_count", "refresh_weighted_mask"):
            assert hasattr(m, attr), (type(m).__name__, attr)
    # state_dict names are the reference's (checkpoints interchange with a plain reference model)
    modules.uninstall()
    plain = ref_shim.build_reference_model(192, 3, depth=2)
    assert list(plain.state_dict()) == list(model.state_dict())
    plain.load_state_dict(model.state_dict())


@needs_ref
def test_weighted_mask_matches_reference_forward(injected):
    L, VT, modules = injected
    model = ref_shim.build_reference_model(192, 3, depth=2)
    g = torch.Generator().manual_seed(1)
    cases = [(model.patch_embed, L.MAEPatchEmbed, (torch.randn(2, 3, 224, 224, generator=g),)),
             (model.blocks[0].attn, L.MAESparseAttention, (torch.randn(2, 197, 192, generator=g),)),
             (model.blocks[1].mlp, L.MAESparseMlp, (torch.randn(2, 197, 192, generator=g),))]
    for m, base, args in cases:
        with torch.no_grad():
            m.switch_cell.view(-1)[1] = False                       # a dead cell
        base.forward(m, *args)                                      # the reference's own forward sets weighted_mask
        ref_wm = m.weighted_mask
        ref_grad, = torch.autograd.grad(ref_wm.square().sum(), m.alpha)
        m.refresh_weighted_mask()
        assert m.weighted_mask.shape == ref_wm.shape
        assert torch.allclose(m.weighted_mask, ref_wm, atol=1e-6)
        got_grad, = torch.autograd.grad(m.weighted_mask.square().sum(), m.alpha)
        assert torch.allclose(got_grad, ref_grad, atol=1e-6)


@needs_ref
def test_no_cpu_fallback_and_pickle(injected):
    L, VT, modules = injected
    from ofb_b200._lib import OfbError
    model = ref_shim.build_reference_model(192, 3, depth=2)
    with pytest.raises(OfbError):
        model(torch.zeros(2, 3, 224, 224))
    with pytest.raises(OfbError):
        model.blocks[0].attn(torch.zeros(2, 197, 192))
    buf = io.BytesIO()
    torch.save(model, buf)                                          # whole-object checkpoint (search.py:671-740)
    buf.seek(0)
    again = torch.load(buf, weights_only=False)
    assert type(again.blocks[0].mlp).__name__ == "OFBSparseMlp"
    assert pickle.loads(pickle.dumps(modules.ModelBridge())).engine is None
