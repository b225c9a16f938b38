"""TEST INFRASTRUCTURE ONLY - drives the UNMODIFIED reference (ref_shim.REFERENCE_ROOT: /root/reference in the build
container, the staged copy oracle/_ref on the GPU box) through the body of engine.search_one_epoch (engine.py:95-184):
model forward, OFBSearchLOSS, decoder-loss weighting, backward, optimizer_param / optimizer_arch / optimizer_decoder step.

Users: bench.py's reference arm (`--impl reference`, `cpu_baseline`, `gpu_eager_baseline`) and the drop-in tests.  The
product never imports this.  Nothing here restates reference arithmetic: the model, criterion and optimizers are the
reference's own classes, built as search.py:393-559 / 583-600 builds them.
"""
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
import ref_shim  # noqa: E402


def build_optimizers(model, lr, wd=1e-3):
    """search.py:486-559: optimizer_param (no-decay | decay), optimizer_decoder (same split), optimizer_arch (alphas)."""
    import optim as ref_optim
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
        groups[key].append(p)
        names[key].append(name)
    opt_param = ref_optim.AdamW([{"params": groups["param_nd"], "weight_decay": 0.},
                                 {"params": groups["param_d"], "weight_decay": wd}],
                                {0: names["param_nd"], 1: names["param_d"]}, lr=lr, eps=1e-8, betas=(0.9, 0.999))
    opt_dec = ref_optim.AdamW([{"params": groups["dec_nd"], "weight_decay": 0.},
                               {"params": groups["dec_d"], "weight_decay": wd}],
                              {0: names["dec_nd"], 1: names["dec_d"]}, lr=lr, eps=1e-8, betas=(0.9, 0.999))
    opt_arch = ref_optim.AdamW(groups["arch"], {0: names["arch"]}, lr=lr, eps=1e-8, betas=(0.5, 0.999), weight_decay=wd)
    return opt_param, opt_dec, opt_arch


def build_criterion(device):
    """search.py:583-600 with the README's loss weights (search.py:173-179 defaults)."""
    from losses import DistillationLoss, OFBSearchLOSS
    return OFBSearchLOSS(DistillationLoss(ref_shim.LabelSmoothingCrossEntropy(0.1), None, "none", 0.5, 1.0), device,
                         attn_w=0.5, mlp_w=0.5, patch_w=0, embedding_w=0.5, flops_w=5, entropy=True, var=True, norm=True)


class ReferenceStep:
    """The reference's search step on `device` ("cpu" or "cuda"); autocast=None (fp32, as the reference runs by default) or a
    torch dtype (torch.bfloat16: PyTorch-eager mixed precision, the same-box GPU yardstick of SURVEY 8d)."""

    def __init__(self, embed_dim, num_heads, depth=12, device="cpu", lr=1e-3, drop_path_rate=0.1, autocast=None,
                 target_flops=1.0, epoch_frac=0.0, seed=0, quiet=True):
        ref_shim.install()
        self.dev = torch.device(device)
        self.model = ref_shim.build_reference_model(embed_dim, num_heads, depth, drop_path_rate, seed=seed).to(self.dev)
        self.model.train()
        self.ddp = ref_shim.FakeDDP(self.model)
        self.opt_param, self.opt_dec, self.opt_arch = build_optimizers(self.model, lr)
        self.criterion = build_criterion(self.dev)
        self.autocast, self.target_flops, self.quiet = autocast, target_flops, quiet
        self.model.adjust_masking_ratio(epoch_frac, 20, 100, max_ratio=0.95, min_ratio=0.75)
        for m in self.model.searchable_modules:
            if not m.finish_search:
                m.update_w(epoch_frac, 20)

    def step(self, images, labels):
        """engine.py:131-184 (non-amp branch; with autocast the forward + criterion run under torch.autocast)."""
        out_fd = None
        if self.quiet:            # get_flops_loss prints a formatted line every step (base_model.py:33)
            sys.stdout.flush()
            out_fd = os.dup(1)
            null = os.open(os.devnull, os.O_WRONLY)
            os.dup2(null, 1)
            os.close(null)
        try:
            ctx = torch.autocast(self.dev.type, dtype=self.autocast) if self.autocast is not None else _Null()
            with ctx:
                outputs, (dec_loss, score_loss) = self.ddp(images)
                base, arch = self.criterion(images, outputs, labels, self.ddp, "arch", self.target_flops, False)
            total = base + arch
            if dec_loss != 0.:
                total = total + (base / dec_loss).data.clone() * dec_loss
            loss_value = total.item()
            total.backward()
            self.opt_param.step(); self.opt_arch.step(); self.opt_dec.step()
            self.opt_param.zero_grad(); self.opt_arch.zero_grad(); self.opt_dec.zero_grad()
        finally:
            if out_fd is not None:
                sys.stdout.flush()
                os.dup2(out_fd, 1)
                os.close(out_fd)
        return loss_value


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def time_reference(embed_dim, num_heads, depth, batch, device="cpu", steps=3, warmup=1, autocast=None, threads=None, seed=1):
    """(images/s, seconds/step, cores) of the unmodified reference step on synthetic ImageNet-shaped input."""
    cores = threads or os.cpu_count() or 1
    if torch.device(device).type == "cpu":
        torch.set_num_threads(cores)
    rs = ReferenceStep(embed_dim, num_heads, depth, device=device, autocast=autocast)
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(batch, 3, 224, 224, generator=g).to(rs.dev)
    labels = torch.randint(0, 1000, (batch,), generator=g).to(rs.dev)
    sync = (lambda: torch.cuda.synchronize(rs.dev)) if rs.dev.type == "cuda" else (lambda: None)
    for _ in range(warmup):
        rs.step(images, labels)
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        rs.step(images, labels)
    sync()
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt, cores
