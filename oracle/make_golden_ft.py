"""TEST INFRASTRUCTURE ONLY — runs the UNMODIFIED reference classes (/root/reference via ref_shim) for one *finetune*
training step of a physically pruned subnet, checks oracle/ft_oracle.py against it and writes tests/golden/ft_<case>.npz.

Run in the build container only:   python oracle/make_golden_ft.py
Reference side: plain VisionTransformer (models/vision_transformer.py:226-358) whose layers are replaced by pruned-shape
parameters exactly the way finetune.intersect does it (finetune.py:182-249: new nn.Parameter per tensor, in/out_features,
LayerNorm.normalized_shape[0], Attention.num_heads), lr_decay.param_groups_lrd + torch.optim.AdamW (finetune.py:378-383),
criterion = DistillationLoss(LabelSmoothingCrossEntropy | SoftTargetCrossEntropy, None, 'none') (finetune.py:388-415),
then the body of engine.train_one_epoch (engine.py:31-62).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from fixtures import summarize  # noqa: E402
from ft_oracle import SubnetCfg, ft_group, ft_train_step, make_ft_inputs, make_ft_params, torch_adamw_step  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")

CASES = {
    # pruned DeiT-S-like subnets (embed widths are multiples of 12, head dims multiples of 8, hidden multiples of 192)
    "ft_s_d3_b2": dict(base=(384, 6), D=252, heads=[4, 6, 2], head_dims=[40, 64, 56], hiddens=[768, 1152, 384], B=2,
                       train=False, dpr=0.0, soft=False, lr=1e-3),
    "ft_s_d2_b3_soft_dp": dict(base=(384, 6), D=336, heads=[6, 4], head_dims=[48, 32], hiddens=[960, 1536], B=3,
                               train=True, dpr=0.1, soft=True, lr=5e-4),
    "ft_t_d12_b2": dict(base=(192, 3), D=168, heads=[2] * 12, head_dims=[64, 56, 48, 40, 64, 32, 24, 16, 64, 56, 48, 40],
                        hiddens=[480, 576, 768, 192, 288, 384, 672, 768, 480, 576, 768, 192], B=2, train=False, dpr=0.0,
                        soft=False, lr=1e-3),
}


def build_reference(cfg, base, P0, drop_path_rate):
    ref_shim.install()
    from functools import partial
    from models.layers import LayerNorm, PatchEmbed
    from models.vision_transformer import VisionTransformer
    D0, H0 = base
    model = VisionTransformer(patch_size=16, embed_dim=D0, depth=cfg.depth, num_heads=H0, mlp_ratio=4, qkv_bias=True,
                              norm_layer=partial(LayerNorm, eps=1e-6), embed_layer=PatchEmbed,
                              drop_path_rate=drop_path_rate, num_classes=cfg.num_classes)
    mods = dict(model.named_modules())
    for k, v in P0.items():                       # finetune.intersect, one tensor at a time
        if "." in k:
            mname, attr = k.rsplit(".", 1)
            layer = mods[mname]
            setattr(layer, attr, torch.nn.Parameter(v.clone()))
            if attr == "weight":
                if hasattr(layer, "out_channels"):
                    layer.out_channels, layer.in_channels = v.shape[0], v.shape[1]
                if hasattr(layer, "out_features"):
                    layer.out_features, layer.in_features = v.shape[0], v.shape[1]
                if isinstance(layer, LayerNorm):
                    layer.normalized_shape[0] = v.shape[-1]
        else:
            setattr(model, k, torch.nn.Parameter(v.clone()))
    for l, blk in enumerate(model.blocks):
        blk.attn.num_heads = cfg.heads[l]         # finetune.py:238-242
        assert abs(blk.attn.scale - cfg.scale) < 1e-12
    assert [k for k, _ in model.named_parameters()] == list(P0), "parameter order differs from the reference"
    return model


def run_reference(cfg, c, P0, images, labels, drop_scale, target):
    ref_shim.install()
    import lr_decay as lrd
    from losses import DistillationLoss
    model = build_reference(cfg, c["base"], P0, c["dpr"])
    groups = lrd.param_groups_lrd(model, 0.05, no_weight_decay_list=model.no_weight_decay(), layer_decay=0.95)
    opt = torch.optim.AdamW(groups, lr=c["lr"], eps=1e-8)
    for gr in opt.param_groups:                   # what the timm scheduler applies every update (lr * lr_scale)
        gr["lr"] = c["lr"] * gr["lr_scale"]
    base_crit = ref_shim.SoftTargetCrossEntropy() if c["soft"] else ref_shim.LabelSmoothingCrossEntropy(0.1)
    criterion = DistillationLoss(base_crit, None, "none", 0.5, 1.0)
    model.train(c["train"])                       # finetune.py:445: eval mode when finetuning from a checkpoint
    queue = []
    if c["train"]:
        dpr = torch.linspace(0, c["dpr"], cfg.depth)
        for l in range(cfg.depth):
            if float(dpr[l]) > 0:
                keep = 1 - float(dpr[l])
                # invert floor(keep + u) / keep: any u reproducing the stored multiplier
                for j in range(2):
                    queue.append(torch.where(drop_scale[l, j] > 0, torch.full_like(drop_scale[l, j], 0.999),
                                             torch.zeros_like(drop_scale[l, j])).reshape(-1, 1, 1) * (1 - keep) / (1 - keep)
                                 if keep < 1 else None)
    real_rand = torch.rand

    def fake_rand(*a, **k):
        t = queue.pop(0)
        return t.clone()

    torch.rand = fake_rand
    try:
        opt.zero_grad()
        outputs = model(images.clone())
        loss = criterion(images, outputs, target if c["soft"] else labels)
    finally:
        torch.rand = real_rand
    assert not queue
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    opt.step()
    new_params = {k: p.detach().clone() for k, p in model.named_parameters()}
    return dict(logits=outputs.detach(), loss=loss.detach(), grads=grads, new_params=new_params)


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    for name, c in CASES.items():
        cfg = SubnetCfg(embed_dim=c["D"], heads=c["heads"], head_dims=c["head_dims"], hiddens=c["hiddens"])
        P0 = make_ft_params(cfg, seed=0)
        images, labels, drop_scale, target = make_ft_inputs(cfg, c["B"], seed=1, drop_path_rate=c["dpr"] if c["train"] else 0.0,
                                                            soft=c["soft"])
        ref = run_reference(cfg, c, P0, images, labels, drop_scale, target)
        P = {k: v.clone() for k, v in P0.items()}
        logits, loss, grads = ft_train_step(P, {}, images, labels, cfg, lr=c["lr"], step=1, drop_scale=drop_scale,
                                            target=target, update=False)
        checks = {"logits": rel(logits, ref["logits"]), "loss": rel(loss, ref["loss"])}
        for k, g in ref["grads"].items():
            checks["grad:" + k] = rel(grads[k], g)
            # the AdamW restatement is pinned on the reference's own gradient (step 1 of Adam is ~lr * sign(g))
            _, sc, wd = ft_group(k, P0[k].shape, cfg.depth, 0.05, 0.95)
            pk, mk, vk = P0[k].clone(), torch.zeros_like(P0[k]), torch.zeros_like(P0[k])
            torch_adamw_step(pk, g, mk, vk, 1, c["lr"] * sc, wd)
            checks["new:" + k] = rel(pk, ref["new_params"][k])
        bad = {k: v for k, v in checks.items() if v > (1e-6 if k.startswith("new:") else 1e-4)}
        print(f"[{name}] oracle vs reference: worst rel err {max(checks.values()):.3e} over {len(checks)} tensors; "
              f"loss={float(ref['loss']):.6f}")
        assert not bad, bad
        gold = {"logits": ref["logits"].numpy(), "loss": ref["loss"].numpy(), "D": np.array(c["D"]),
                "base": np.array(c["base"]), "heads": np.array(c["heads"]), "head_dims": np.array(c["head_dims"]),
                "hiddens": np.array(c["hiddens"]), "B": np.array(c["B"]), "train": np.array(c["train"]),
                "dpr": np.array(c["dpr"]), "soft": np.array(c["soft"]), "lr": np.array(c["lr"])}
        for k, g in ref["grads"].items():
            gold["gsum:" + k] = summarize(g).numpy()
            gold["psum:" + k] = summarize(ref["new_params"][k]).numpy()
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **gold)
        print(f"  wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
