"""TEST INFRASTRUCTURE — run one finetune training step of a pruned subnet on the GPU engine (FinetuneStepEngine) and on the
CPU oracle (oracle/ft_oracle.py) with identical seeded parameters / inputs, compare logits, loss, every gradient, the AdamW
update (layer-decay groups) and that the zero padding of the pruned layout stayed exactly zero.
Tolerances as in step_compare.py (bf16 compute vs fp32 oracle: rel 2e-2; pure-fp32 pieces 1e-4)."""
import torch

from step_compare import BF16_TOL, FP32_TOL, GRAD_MAX_TOL, LOSS_TOL, rel, rel_l2


def compare_ft_step(embed_dim, heads, head_dims, hiddens, batch, lr=1e-3, train=False, dpr=0.0, soft=False, verbose=False):
    import ofb_b200  # noqa: F401
    from ft_oracle import SubnetCfg, ft_group, ft_train_step, make_ft_inputs, make_ft_params, torch_adamw_step
    from ofb_b200.finetune_engine import FinetuneStepEngine

    cfg = SubnetCfg(embed_dim=embed_dim, heads=heads, head_dims=head_dims, hiddens=hiddens)
    P = make_ft_params(cfg, seed=0)
    images, labels, drop_scale, target = make_ft_inputs(cfg, batch, seed=1, drop_path_rate=dpr if train else 0.0, soft=soft)
    eng = FinetuneStepEngine(embed_dim, heads, head_dims, hiddens, batch, lr=lr, training_mode=train, drop_path_rate=dpr)
    eng.load_params(P)
    drop_u = (drop_scale > 0).float().reshape(cfg.depth * 2, batch) * 0.999
    scal = eng.step(images.cuda(), labels.cuda(), target.cuda() if target is not None else None, drop_u=drop_u.cuda(),
                    update=False)
    torch.cuda.synchronize()
    logits, loss, grads = ft_train_step({k: v.clone() for k, v in P.items()}, {}, images, labels, cfg, lr=lr, step=1,
                                        drop_scale=drop_scale, target=target, update=False)
    errs = {"logits": rel(eng.logits, logits), "loss": rel(scal[0], loss)}
    got = eng.named_grads()
    gerrs = {k: rel_l2(got[k], g) for k, g in grads.items()}
    gmax = {k: rel(got[k], g) for k, g in grads.items()}
    clean_before = eng.padding_is_clean()
    # AdamW: the fused kernel on the engine's own gradients vs torch.optim.AdamW semantics on the same gradients
    g_engine = {k: v.detach().cpu().clone() for k, v in got.items()}
    p_before = {k: v.detach().cpu().clone() for k, v in eng.named_parameters().items()}
    eng.optimizer_step()
    torch.cuda.synchronize()
    after = eng.named_parameters()
    aerrs = {}
    for k in p_before:
        _, sc, wd = ft_group(k, p_before[k].shape, cfg.depth, 0.05, 0.95)
        pk, mk, vk = p_before[k].clone(), torch.zeros_like(p_before[k]), torch.zeros_like(p_before[k])
        torch_adamw_step(pk, g_engine[k], mk, vk, 1, lr * sc, wd)
        aerrs[k] = rel(after[k], pk)
    clean_after = eng.padding_is_clean() and float(eng.grads.abs().max()) == 0.0
    worst_g = max(gerrs.items(), key=lambda kv: kv[1])
    worst_m = max(gmax.items(), key=lambda kv: kv[1])
    worst_a = max(aerrs.items(), key=lambda kv: kv[1])
    ok = (errs["logits"] < BF16_TOL and errs["loss"] < LOSS_TOL and worst_g[1] < BF16_TOL and worst_m[1] < GRAD_MAX_TOL
          and worst_a[1] < FP32_TOL and clean_before and clean_after)
    summary = (f"ft D{embed_dim} heads{heads} d{head_dims} hid{hiddens} B{batch}: logits {errs['logits']:.2e} loss "
               f"{errs['loss']:.2e} worst-grad(L2) {worst_g[0]} {worst_g[1]:.2e} worst-grad(max) {worst_m[0]} {worst_m[1]:.2e} "
               f"worst-adamw {worst_a[0]} {worst_a[1]:.2e} padding clean {clean_before}/{clean_after} ok={ok}")
    if verbose:
        for k, v in sorted(gmax.items(), key=lambda kv: -kv[1])[:8]:
            print(f"   grad {k}: max-rel {v:.3e}  l2-rel {gerrs[k]:.3e}")
    return dict(ok=ok, summary=summary, errs=errs, grad_errs=gerrs, adamw_errs=aerrs, eng=eng)
