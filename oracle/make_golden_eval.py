"""TEST INFRASTRUCTURE ONLY — runs the UNMODIFIED reference (/root/reference via ref_shim) in EVAL mode on an unfinished
search model, exactly as engine.evaluate does (engine.py:222-257: model.eval(); output, _ = model(images);
CrossEntropyLoss; timm accuracy top-1/top-5), checks the oracle restatement (forward_step with no PMIM masking and identity
DropPath) against it and writes tests/golden/eval/<case>.npz.   Run in the build container only.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from fixtures import make_inputs, make_params  # noqa: E402
from make_golden import kill_cells  # noqa: E402
from ofb_oracle import ModelCfg, default_switches, forward_step  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden", "eval")
CASES = {
    "tiny_d3_b8_e10": dict(D=192, H=3, depth=3, B=8, epoch_frac=10.0, dead=False),
    "small_d2_b6_e5_dead": dict(D=384, H=6, depth=2, B=6, epoch_frac=5.0, dead=True),
}


def eval_inputs(cfg, c):
    """The step fixtures with the eval-mode settings: every patch kept, DropPath multipliers 1."""
    inp = make_inputs(cfg, c["B"], seed=1, epoch_frac=c["epoch_frac"], drop_path_rate=0.0, keep_ratio=1.0)
    return inp


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    ref_shim.install()
    for name, c in CASES.items():
        cfg = ModelCfg(embed_dim=c["D"], num_heads=c["H"], depth=c["depth"])
        P0 = make_params(cfg, seed=0)
        # make the head informative so that top-1 / top-5 are not all misses: bias the label logit through a label-dependent
        # shift of the head bias is not possible (labels are inputs) - instead keep random logits and store exact hit flags
        inp = eval_inputs(cfg, c)
        switches = default_switches(cfg)
        if c["dead"]:
            switches = kill_cells(switches)
        model = ref_shim.build_reference_model(cfg.embed_dim, cfg.num_heads, cfg.depth, 0.1, cfg.num_classes)
        with torch.no_grad():
            for k, p in model.named_parameters():
                p.copy_(P0[k])
        mods = {"patch_embed": model.patch_embed}
        for l, blk in enumerate(model.blocks):
            mods[f"blocks.{l}.attn"] = blk.attn
            mods[f"blocks.{l}.mlp"] = blk.mlp
        for k, m in mods.items():
            m.switch_cell = switches[k].clone()
        for m in model.searchable_modules:
            m.update_w(c["epoch_frac"], 20)
        model.eval()
        # labels: half of them set to the reference's own arg-max / 3rd-best class so that hits and misses both occur
        with torch.no_grad():
            output, _ = model(inp.images.clone())
            order = output.argsort(dim=1, descending=True)
            labels = inp.labels.clone()
            labels[0::4] = order[0::4, 0]
            labels[1::4] = order[1::4, 2]
            labels[2::4] = order[2::4, 7]
            loss = torch.nn.CrossEntropyLoss()(output, labels)
            from timm.utils import accuracy
            acc1, acc5 = accuracy(output, labels, topk=(1, 5))
        out = forward_step({k: v.clone() for k, v in P0.items()}, inp, cfg, switches)
        err = float((out.logits - output).abs().max() / output.abs().max())
        print(f"[{name}] oracle vs reference eval logits: rel {err:.3e}; loss {float(loss):.5f} acc1 {float(acc1):.2f} acc5 {float(acc5):.2f}")
        assert err < 1e-5
        gold = {"logits": output.numpy(), "labels": labels.numpy(), "loss": loss.numpy(), "acc1": acc1.numpy(),
                "acc5": acc5.numpy(), "cfg": np.array([c["D"], c["H"], c["depth"], c["B"]]),
                "epoch_frac": np.array(c["epoch_frac"])}
        for k, s in switches.items():
            gold["switch:" + k] = s.numpy()
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **gold)


if __name__ == "__main__":
    main()
