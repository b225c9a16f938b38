"""TEST INFRASTRUCTURE ONLY — runs the UNMODIFIED reference (/root/reference, via ref_shim) for one full search step
on seeded parameters/inputs, checks the oracle restatement against it and writes tests/golden/<case>.npz.

Run in the build container only (the reference does not travel to the GPU box):
    python oracle/make_golden.py
The step executed on the reference side is the body of engine.search_one_epoch (engine.py:102-184) with the optimizers
built as in search.py:486-559 (reference optim.AdamW).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from fixtures import make_inputs, make_params, summarize  # noqa: E402
from ofb_oracle import ModelCfg, adamw_step, default_switches, group_hparams, param_group, train_step  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")

CASES = {
    # name: (embed_dim, heads, depth, batch, epoch_frac, drop_path, dead_cells, lr)
    # batch >= 8: the fixed bf16 gradient bound of tests/step_compare.py is asserted against these fixtures (mask_token and
    # the decoder gradients are sums over ~10 masked tokens per image: at batch 2-3 they are too few to average anything)
    "tiny_d12_b16_e0": dict(D=192, H=3, depth=12, B=16, epoch_frac=0.0, dpr=0.1, dead=False, lr=1e-3),
    "tiny_d3_b8_e10": dict(D=192, H=3, depth=3, B=8, epoch_frac=10.0, dpr=0.0, dead=False, lr=1e-3),
    "small_d2_b8_e5_dead": dict(D=384, H=6, depth=2, B=8, epoch_frac=5.0, dpr=0.1, dead=True, lr=1e-3),
}


def kill_cells(switches, seed=7):
    g = torch.Generator().manual_seed(seed)
    for k, s in switches.items():
        dead = torch.rand(s.shape, generator=g) < 0.3
        dead.view(-1)[int(torch.randint(0, s.numel(), (1,), generator=g))] = False   # keep >= 1 alive
        if (~dead).sum() < 2:
            dead[:] = False
        switches[k] = ~dead
    return switches


def run_reference(cfg, P0, inp, switches, drop_path_rate, lr, epoch_frac, after_load=None):
    """after_load(model): optional hook run once the parameters and switch cells are in place and before the optimizers are
    built (make_golden_pruned_step.py uses it to run compress())."""
    ref_shim.install()
    import optim as ref_optim
    from losses import DistillationLoss, OFBSearchLOSS
    model = ref_shim.build_reference_model(cfg.embed_dim, cfg.num_heads, cfg.depth, drop_path_rate, cfg.num_classes)
    with torch.no_grad():
        sd = dict(model.named_parameters())
        assert set(sd) == set(P0), (set(sd) ^ set(P0))
        for k, p in sd.items():
            p.copy_(P0[k])
    mods = {"patch_embed": model.patch_embed}
    for l, blk in enumerate(model.blocks):
        mods[f"blocks.{l}.attn"] = blk.attn
        mods[f"blocks.{l}.mlp"] = blk.mlp
    for k, m in mods.items():
        m.switch_cell = switches[k].clone()
    if after_load is not None:
        after_load(model)
    model.train()
    ddp = ref_shim.FakeDDP(model)

    # optimizers: search.py:486-559
    groups = {"param_nd": [], "param_d": [], "dec_nd": [], "dec_d": [], "arch": []}
    names = {k: [] for k in groups}
    skip = model.no_weight_decay()
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        if len(p.shape) == 1 or name.endswith(".bias") or any(e in name for e in skip):
            key = "param_nd" if "decoder" not in name else "dec_nd"
        elif "alpha" in name:
            key = "arch"
        else:
            key = "param_d" if "decoder" not in name else "dec_d"
        assert key == param_group(name, p), name
        groups[key].append(p)
        names[key].append(name)
    opt_param = ref_optim.AdamW([{"params": groups["param_nd"], "weight_decay": 0.},
                                 {"params": groups["param_d"], "weight_decay": 1e-3}],
                                {0: names["param_nd"], 1: names["param_d"]}, lr=lr, eps=1e-8, betas=(0.9, 0.999))
    opt_dec = ref_optim.AdamW([{"params": groups["dec_nd"], "weight_decay": 0.},
                               {"params": groups["dec_d"], "weight_decay": 1e-3}],
                              {0: names["dec_nd"], 1: names["dec_d"]}, lr=lr, eps=1e-8, betas=(0.9, 0.999))
    opt_arch = ref_optim.AdamW(groups["arch"], {0: names["arch"]}, lr=lr, eps=1e-8, betas=(0.5, 0.999),
                               weight_decay=1e-3)
    criterion = OFBSearchLOSS(
        DistillationLoss(ref_shim.LabelSmoothingCrossEntropy(0.1), None, "none", 0.5, 1.0), torch.device("cpu"),
        attn_w=0.5, mlp_w=0.5, patch_w=0, embedding_w=0.5, flops_w=5, entropy=True, var=True, norm=True)

    # engine.py:102-117
    model.adjust_masking_ratio(epoch_frac, 20, 100, max_ratio=0.95, min_ratio=0.75)
    for m in model.searchable_modules:
        if not m.finish_search:
            m.update_w(epoch_frac, 20)

    # feed the recorded random draws to the reference's torch.rand calls (PMIM noise, then DropPath per block)
    queue = [inp.noise] + [u.reshape(-1, 1, 1) for u in inp.drop_draws]
    real_rand = torch.rand

    def fake_rand(*a, **k):
        t = queue.pop(0)
        shape = tuple(a[0]) if len(a) == 1 and isinstance(a[0], (tuple, list, torch.Size)) else tuple(a)
        assert tuple(t.shape) == shape, (t.shape, shape)
        return t.clone()

    torch.rand = fake_rand
    try:
        outputs, (dec_loss, score_loss) = ddp(inp.images.clone())
    finally:
        torch.rand = real_rand
    assert not queue and score_loss is None
    gates = {}
    with torch.no_grad():
        for k, m in mods.items():
            if "mlp" in k:   # MAESparseMlp.get_weight returns nothing (layers.py:877-881): redo its two lines
                rank = torch.argsort(torch.argsort(m.score, dim=-1, descending=True), dim=-1)
                wr, ps = torch.gather(m.weighted_mask, -1, rank), m.score.sigmoid()
            else:
                wr, ps = m.get_weight()
            gates[k] = (m.w_p * ps + (1 - m.w_p) * wr).reshape(-1).clone()
    base, arch = criterion(inp.images, outputs, inp.labels, ddp, "arch", cfg.target_flops, False)
    total = base + arch
    w_dec = (base / dec_loss).data.clone()
    total = total + w_dec * dec_loss
    total.backward()
    grads = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in model.named_parameters()}
    opt_param.step(); opt_arch.step(); opt_dec.step()
    new_params = {k: p.detach().clone() for k, p in model.named_parameters()}
    return dict(logits=outputs.detach(), base=base.detach(), arch=arch.detach(), dec=dec_loss.detach(),
                total=total.detach(), grads=grads, new_params=new_params, gates=gates,
                flops=[float(x) for x in model.get_flops()])


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    for name, c in CASES.items():
        cfg = ModelCfg(embed_dim=c["D"], num_heads=c["H"], depth=c["depth"])
        P0 = make_params(cfg, seed=0)
        inp = make_inputs(cfg, c["B"], seed=1, epoch_frac=c["epoch_frac"], drop_path_rate=c["dpr"])
        switches = default_switches(cfg)
        if c["dead"]:
            switches = kill_cells(switches)
        ref = run_reference(cfg, P0, inp, switches, c["dpr"], c["lr"], c["epoch_frac"])

        P = {k: v.clone() for k, v in P0.items()}
        out, grads = train_step(P, {}, inp, cfg, lr=c["lr"], step=1, switches=switches)
        worst = 0.
        checks = {"logits": rel(out.logits, ref["logits"]), "base": rel(out.loss_base, ref["base"]),
                  "arch": rel(out.loss_arch, ref["arch"]), "dec": rel(out.loss_decoder, ref["dec"]),
                  "total": rel(out.loss_total, ref["total"])}
        for k, g in ref["grads"].items():
            if g is None:
                assert grads.get(k) is None, k
                continue
            checks["grad:" + k] = rel(grads[k], g)
            # AdamW restatement is pinned on the reference's own gradient (step 1 of Adam is ~lr*sign(g), so feeding
            # it the oracle's gradient would amplify 1e-7 noise on near-zero entries)
            pk, mk, vk = P0[k].clone(), torch.zeros_like(P0[k]), torch.zeros_like(P0[k])
            adamw_step(pk, g, mk, vk, 1, **group_hparams(param_group(k, pk), c["lr"]))
            checks["new:" + k] = rel(pk, ref["new_params"][k])
        for k, gt in ref["gates"].items():
            checks["gate:" + k] = rel(out.gates[k].reshape(-1), gt)
        bad = {k: v for k, v in checks.items() if v > (1e-6 if k.startswith('new:') else 1e-4)}
        worst = max(checks.values())
        print(f"[{name}] oracle vs reference: worst rel err {worst:.3e} over {len(checks)} tensors; "
              f"losses base={float(ref['base']):.6f} arch={float(ref['arch']):.6f} dec={float(ref['dec']):.6f} "
              f"flops={ref['flops']}")
        assert not bad, bad

        gold = {"logits": ref["logits"].numpy(), "loss_base": ref["base"].numpy(), "loss_arch": ref["arch"].numpy(),
                "loss_decoder": ref["dec"].numpy(), "loss_total": ref["total"].numpy(),
                "flops": np.array(ref["flops"]),
                "cfg": np.array([c["D"], c["H"], c["depth"], c["B"]]), "epoch_frac": np.array(c["epoch_frac"]),
                "dpr": np.array(c["dpr"]), "lr": np.array(c["lr"]), "dead": np.array(c["dead"])}
        for k, g in ref["grads"].items():
            if g is not None:
                gold["gsum:" + k] = summarize(g).numpy()
                gold["gmax:" + k] = np.array(float(g.abs().max()))      # scale of the sampled-entry comparison
                gold["psum:" + k] = summarize(ref["new_params"][k]).numpy()
        for k, gt in ref["gates"].items():
            gold["gate:" + k] = gt.numpy()
        for k, s in switches.items():
            gold["switch:" + k] = s.numpy()
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **gold)
        print(f"  wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
