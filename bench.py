"""bench.py — images/s of the DeiT-S bi-mask search step (fwd + bwd + 3x AdamW, PMIM on), the BASELINE.json metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--model small|tiny|base] [--batch B]
  torchrun launches it with one rank per GPU for N > 1 (weak scaling: 256 images per GPU, NCCL gradient all-reduce).

One JSON line on rank 0:
  value      whole-job images/s with the batch already resident in HBM (device-timed, max over ranks)
  e2e        same metric through the engine's public step() with HOST (pinned) images/labels: H2D copy of every step's
             inputs and a D2H read of the step's loss vector inside the timed region
  roofline   dominant kernel = ofb::gemm_kernel (tcgen05 GEMM, all instantiations): achieved = algorithmic FLOPs of the
             GEMM launches of the timed steps / their CUDA-event durations, against the measured bf16 peak
  cpu_baseline / --impl reference: the CPU oracle port of the reference step (oracle/ofb_oracle.py) on the host cores,
             bounded sample of the same workload (the Python reference itself cannot travel to the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODELS = {"tiny": (192, 3), "small": (384, 6), "base": (768, 12)}
METRIC = {"search": "images/sec DeiT-S bi-mask search step (fwd+bwd+update, PMIM on)",
          "finetune": "images/sec finetune step of a physically pruned DeiT-S subnet (fwd+bwd+update)",
          "post": "images/sec post-search step of the finalised DeiT-S subnet (Mixup/CutMix + soft-target CE, PMIM off, fwd+bwd+update)"}
# algorithmic GFLOP / image of one search step: tiny 7.64, small 27.82, base 105.85
STEP_GFLOP = {"tiny": 7.64, "small": 27.82, "base": 105.85}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ofb", choices=["ofb", "reference"])
    ap.add_argument("--model", default="small", choices=list(MODELS))
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--depth", type=int, default=12)
    ap.add_argument("--cpu-sample-batch", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying CUDA graphs")
    ap.add_argument("--no-eager", action="store_true", help="skip the same-box GPU baseline (unmodified reference, PyTorch eager)")
    ap.add_argument("--no-extra", action="store_true", help="skip the short runs of the other BASELINE.json configurations")
    ap.add_argument("--workload", default="search", choices=["search", "finetune", "post"],
                    help="search: the bi-mask search step (BASELINE.json metric, default); finetune: the training step of a "
                         "physically pruned DeiT-S subnet (BASELINE.json configs[4], extra configuration); post: the post-search "
                         "phase of the search loop (SURVEY 8f-3) on the same subnet shapes, reached through prune_event()")
    return ap.parse_args()


def gemm_traffic():
    """Measured DRAM traffic of the GEMM family per launch (ncu --set full capture of this workload, profiles/): only valid
    for the configuration it was captured on (DeiT-small, batch 256)."""
    path = os.path.join(ROOT, "profiles", "r02c_gemm_traffic.json")
    if os.path.exists(path):
        return json.load(open(path))
    return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(tflops=float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1590.0))), hbm=float(p.get("hbm_gbs", 6650.0)),
                    src="measured (MEASURED_PEAKS.json, sustained)")
    return dict(tflops=1590.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        # NVML from a thread (10 ms period: the default timed region is ~0.3 s); nvidia-smi -lms as the fallback
        self.stop_flag = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.thr = threading.Thread(target=self._poll, daemon=True)
            self.thr.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        bits = [("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)]
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                row = [str(self.index), str(sm), str(mx), "", ""] + ["Active" if (r & b) else "Not Active" for _, b in bits]
                self.rows.append(row)
            except Exception:
                pass
            time.sleep(0.01)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if getattr(self, "nv", None) is not None:
            self.stop_flag = True
            self.thr.join(timeout=1.0)
        elif self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            time.sleep(0.25)
            self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) > 8 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) > 8 and r[2].isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) > 8:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "sm_mhz_min": sm[0] if sm else None}


# BASELINE.json configs[4]: a searched OFB-DeiT-C-like subnet (README.md:23: 1.7 GFLOPs). The release checkpoints are not
# shipped, so the per-layer dims are synthesised (fixed): pruned embedding 288, per-block heads / head dims / hidden widths
# drawn from the DeiT-S search space (heads 2-6, head dims 16-64 step 8, hidden 384-1536 step 192).
FT_SUBNET = dict(embed_dim=288,
                 heads=[6, 4, 4, 6, 4, 4, 6, 4, 4, 4, 4, 6],
                 head_dims=[48, 56, 40, 48, 64, 40, 48, 56, 40, 48, 32, 64],
                 hiddens=[768, 576, 576, 768, 576, 576, 768, 576, 384, 576, 768, 960])


def cpu_ft_oracle_rate(sample_batch, steps=1, warmup=1):
    """images/s of the CPU finetune oracle (reference step restated, fp32, all host threads) on a bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    from ft_oracle import SubnetCfg, ft_train_step, make_ft_inputs, make_ft_params
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = SubnetCfg(**FT_SUBNET)
    P = make_ft_params(cfg, seed=0)
    images, labels, _, _ = make_ft_inputs(cfg, sample_batch, seed=1)
    state = {}
    for i in range(warmup):
        ft_train_step(P, state, images, labels, cfg, lr=1e-3, step=i + 1)
    t0 = time.perf_counter()
    for i in range(steps):
        ft_train_step(P, state, images, labels, cfg, lr=1e-3, step=warmup + i + 1)
    dt = (time.perf_counter() - t0) / steps
    return sample_batch / dt, cores, dt


def script_subnet_alphas(named, depth):
    """Alphas that leave exactly the FT_SUBNET cell of every searchable module alive at the next prune event (DeiT-S search
    space, layers.py:143-152, 425-462, 813-821): the synthetic stand-in for a finished search."""
    import torch
    D, d, hid = 384, 64, 1536
    ew = [int((i / D) * D) for i in range(D // 2, D + 1, min(D // 32, 12))]
    hc = list(range(2, 6 + 1, 2))
    cw = [int(d * (i / d)) for i in range(d // 4, d + 1, max(d // 8, 1))]
    hw = [int((i / hid) * hid) for i in range(hid // 4, hid + 1, hid // 8)]

    def one(shape, idx):
        a = torch.full(shape, -9.0)
        a.view(-1)[idx] = 2.0
        return a
    out = dict(named)
    out["patch_embed.alpha"] = one((1, len(ew)), ew.index(FT_SUBNET["embed_dim"]))
    for l in range(depth):
        out[f"blocks.{l}.attn.alpha"] = one((len(hc), len(cw)), hc.index(FT_SUBNET["heads"][l]) * len(cw)
                                            + cw.index(FT_SUBNET["head_dims"][l]))
        out[f"blocks.{l}.mlp.alpha"] = one((1, len(hw)), hw.index(FT_SUBNET["hiddens"][l]))
    return out


def subnet_step_gflop():
    N, L, D, C = 197, 196, FT_SUBNET["embed_dim"], 1000
    f = 2.0 * L * 768 * D + 2.0 * D * C
    for H, d, hid in zip(FT_SUBNET["heads"], FT_SUBNET["head_dims"], FT_SUBNET["hiddens"]):
        f += 2.0 * N * D * 3 * H * d + 2.0 * N * H * d * D + 4.0 * N * D * hid + 4.0 * H * N * N * d
    return (3.0 * f - 2.0 * L * 768 * D) / 1e9


def cpu_post_oracle_rate(sample_batch, steps=1, warmup=1):
    """images/s of the CPU oracle on the post-search step (finalised subnet, Mixup soft targets, PMIM off, decoder frozen)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    from fixtures import make_inputs, make_params, pruned_shape_from_plans
    from make_golden_post import FROZEN_P2
    from make_golden_pruned_step import plan_on_cpu
    from ofb_oracle import ModelCfg, default_switches, mixup_batch, train_step
    import ofb_b200  # noqa: F401
    from ofb_b200 import prune
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = ModelCfg(embed_dim=384, num_heads=6, depth=12)
    P0 = script_subnet_alphas(make_params(cfg, seed=0), 12)
    g = torch.Generator().manual_seed(3)
    for k in sorted(k for k in P0 if k.endswith(".score")):
        P0[k] = torch.randn(P0[k].shape, generator=g) * 0.2          # no ties (see oracle/make_golden_prune.py)
    plans, dims = plan_on_cpu(cfg, P0, default_switches(cfg))
    Pp = prune.gather_pruned(plans, {k: v for k, v in P0.items() if k != "alpha_patch"}, dims, 0.545)
    Pp["alpha_patch"] = P0["alpha_patch"]
    shape = pruned_shape_from_plans(cfg, plans)
    sw = {k: pl.switch for k, pl in plans.items()}
    inp = make_inputs(cfg, sample_batch, seed=1, epoch_frac=30.0, drop_path_rate=0.1, keep_ratio=1.0)
    inp.images, inp.soft_target = mixup_batch(inp.images, inp.labels, 0.37, None)
    state = {}
    for i in range(warmup):
        train_step(Pp, state, inp, cfg, lr=1e-3, step=i + 1, switches=sw, shape=shape, frozen=FROZEN_P2, finish_search=True)
    t0 = time.perf_counter()
    for i in range(steps):
        train_step(Pp, state, inp, cfg, lr=1e-3, step=warmup + i + 1, switches=sw, shape=shape, frozen=FROZEN_P2,
                   finish_search=True)
    dt = (time.perf_counter() - t0) / steps
    return sample_batch / dt, cores, dt


def cpu_oracle_rate(model, depth, sample_batch, steps=1, warmup=1):
    """images/s of the CPU oracle port (reference step restated, fp32, all host threads) on a bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    from fixtures import make_inputs, make_params
    from ofb_oracle import ModelCfg, train_step
    D, H = MODELS[model]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = ModelCfg(embed_dim=D, num_heads=H, depth=depth)
    P = make_params(cfg, seed=0)
    inp = make_inputs(cfg, sample_batch, seed=1, epoch_frac=0.0, drop_path_rate=0.1)
    state = {}
    for i in range(warmup):
        train_step(P, state, inp, cfg, lr=1e-3, step=i + 1)
    t0 = time.perf_counter()
    for i in range(steps):
        train_step(P, state, inp, cfg, lr=1e-3, step=warmup + i + 1)
    dt = (time.perf_counter() - t0) / steps
    return sample_batch / dt, cores, dt


def reference_available():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import ref_shim
        return ref_shim.available()
    except Exception:
        return False


def cpu_reference_rate(model, depth, sample_batch, steps=1, warmup=1):
    """images/s of the UNMODIFIED reference step (oracle/_ref or /root/reference through oracle/ref_runner.py: the reference's
    own model, criterion and optimizers; fp32, all host threads) on a bounded sample of the workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_runner
    D, H = MODELS[model]
    rate, dt, cores = ref_runner.time_reference(D, H, depth, sample_batch, device="cpu", steps=steps, warmup=warmup)
    return rate, cores, dt


def cpu_arm(args, steps=1, warmup=1):
    """(rate, cores, s/step, kind) of the CPU arm: the unmodified reference when it is staged (search workload), else the
    oracle port."""
    if args.workload == "finetune":
        return cpu_ft_oracle_rate(args.cpu_sample_batch, steps, warmup) + ("port",)
    if args.workload == "post":
        return cpu_post_oracle_rate(args.cpu_sample_batch, steps, warmup) + ("port",)
    if reference_available():
        return cpu_reference_rate(args.model, args.depth, args.cpu_sample_batch, steps, warmup) + ("reference",)
    return cpu_oracle_rate(args.model, args.depth, args.cpu_sample_batch, steps, warmup) + ("port",)


def gpu_eager_baseline(model, depth, batch, steps=3, warmup=2):
    """The same-box GPU yardstick of SURVEY 2.2 / 8(d): the UNMODIFIED reference (PyTorch eager: cuBLAS / cuDNN / ATen kernels)
    on this B200, fp32 as the reference runs by default and under torch.autocast(bf16)."""
    if not reference_available():
        return {"unavailable": "oracle/_ref not staged (run oracle/make_ref.py in the build container)"}
    import torch
    import ref_runner
    D, H = MODELS[model]
    out = {"kind": "reference (unmodified, PyTorch eager on the same GPU)", "batch": batch, "steps": steps, "warmup": warmup,
           "unit": "images/s"}
    for name, ac in (("fp32", None), ("autocast_bf16", torch.bfloat16)):
        try:
            torch.backends.cuda.matmul.allow_tf32 = False
            rate, dt, _ = ref_runner.time_reference(D, H, depth, batch, device="cuda", steps=steps, warmup=warmup, autocast=ac)
            out[name] = rate
            out[name + "_ms_per_step"] = dt * 1e3
        except Exception as e:                      # noqa: BLE001 - a baseline that cannot run is reported, not fatal
            out[name] = None
            out[name + "_error"] = f"{type(e).__name__}: {str(e)[:200]}"
        torch.cuda.empty_cache()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    rate, cores, dt, kind = cpu_arm(args, steps, warmup)
    what = "unmodified reference (oracle/_ref)" if kind == "reference" else "oracle port"
    sample = (f"{args.cpu_sample_batch} images / step of the same DeiT-{args.model} {args.workload} step, {what}, fp32, "
              f"{cores} threads, {steps} steps after {warmup} warm-up")
    line = {
        "impl": "reference", "metric": METRIC[args.workload], "value": rate,
        "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": (f"DeiT-{args.model} bi-mask search + PMIM step, depth {args.depth}, 224px, CPU sample of "
                                f"{args.cpu_sample_batch} images / step"
                                if args.workload == "search" else
                                "finetune step of a pruned DeiT-S subnet (BASELINE.json configs[4]), 224px, CPU sample"
                                if args.workload == "finetune" else
                                "post-search step of the finalised DeiT-S subnet (SURVEY 8f-3), 224px, CPU sample")},
        "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# kernel families of the step, by substring of the kernel name
FAMILIES = (("gemm", ("gemm_kernel",)), ("attn_bwd", ("attn_bwd",)), ("attn_fwd", ("attn_fwd",)), ("ln", ("ln_fwd", "ln_bwd")),
            ("adamw", ("adamw_kernel",)), ("reduce", ("reduce_partials",)), ("nccl", ("nccl",)))


def family_of(name):
    for fam, keys in FAMILIES:
        if any(k in name for k in keys):
            return fam
    return "other"


def kernel_breakdown(run_steps, n):
    """Per-kernel device durations of `n` steps from CUPTI activity records (torch.profiler): independent of how fast the host
    launches (the timed steps are CUDA-graph replays - an event pair around a host launch would time host gaps instead).
    Returns ({family: ms per step}, {kernel name: (ms per step, launches per step)}) or (None, why)."""
    import torch
    ran = False
    try:
        from torch.profiler import ProfilerActivity, profile
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            ran = True
            run_steps(n)
            torch.cuda.synchronize()
        fam, kern = {}, {}
        for ev in prof.key_averages():
            dt = getattr(ev, "device_time_total", None)
            if dt is None:
                dt = getattr(ev, "cuda_time_total", 0.0)
            if not dt or str(getattr(ev, "device_type", "")).endswith("CPU"):
                continue
            ms = float(dt) / 1e3 / n
            kern[ev.key] = (ms, ev.count / n)
            f = family_of(ev.key)
            fam[f] = fam.get(f, 0.0) + ms
        if not kern:
            return None, "torch.profiler returned no device activity records"
        return fam, kern
    except Exception as e:                      # noqa: BLE001
        if not ran:
            run_steps(n)        # N > 1: the steps hold collectives, every rank must run them exactly once whatever the profiler does
            torch.cuda.synchronize()
        return None, f"{type(e).__name__}: {str(e)[:200]}"


def cublas_same_shapes(gemm_shapes, dev, iters=6):
    """cuBLAS (torch.matmul, bf16, fp32 accumulate) on exactly the GEMM shapes / operand layouts of one step, each timed in
    isolation: ms per step if every GEMM of the step were a plain library GEMM (no fused epilogue: the gate / GELU / residual
    / gradient-reduction passes would come on top). The yardstick for "how good is the tensor main loop on THESE shapes"."""
    import torch
    total, worst = 0.0, None
    for (M, N, K, a_mn, b_mn), count in gemm_shapes.items():
        if M * N * K < 1 << 24:
            continue
        try:
            A = torch.randn((K, M) if a_mn else (M, K), device=dev, dtype=torch.bfloat16)
            B = torch.randn((K, N) if b_mn else (N, K), device=dev, dtype=torch.bfloat16)
            Am, Bm = (A.t() if a_mn else A), (B if b_mn else B.t())
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            for _ in range(2):
                torch.matmul(Am, Bm, out=out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                torch.matmul(Am, Bm, out=out)
            e1.record()
            torch.cuda.synchronize()
            total += e0.elapsed_time(e1) / iters * count
            del A, B, out
        except Exception as e:                      # noqa: BLE001
            worst = f"{type(e).__name__}: {str(e)[:120]}"
    return total, worst


def build_workload(args, dev, pg, world, rank):
    """Engine + step callable + description of one bench workload."""
    import torch
    from ofb_b200.engine import SearchStepEngine
    D, H = MODELS[args.model]
    B = args.batch
    lr = 2.5e-4 * B * world / 256                     # search.py:509-518
    ctx = {}
    if args.workload == "finetune":
        from ofb_b200.finetune_engine import FinetuneStepEngine
        eng = FinetuneStepEngine(batch=B, lr=lr, device=dev, process_group=pg, **FT_SUBNET)
        eng.init_params(seed=0)
        step_gflop = eng.step_flops_per_image() / 1e9
        workload = (f"finetune step of a pruned DeiT-S subnet (embed {FT_SUBNET['embed_dim']}, per-block heads / head dims / "
                    f"hidden widths fixed in bench.py FT_SUBNET), depth 12, batch {B}/GPU, 224px (BASELINE.json configs[4], "
                    "extra configuration)")
    elif args.workload == "post":
        # reach the post-search phase the way a run does: search engine -> the prune event that finalises every module (scripted
        # alphas: the FT_SUBNET cells survive) -> rebuilt engine on the subnet shapes -> enter_post_search()
        assert args.model == "small" and args.depth == 12, "the post workload is defined on the DeiT-S subnet of FT_SUBNET"
        from ofb_b200.mixup import Mixup
        eng0 = SearchStepEngine(D, H, args.depth, B, drop_path_rate=0.1, lr=lr, device=dev, process_group=pg)
        eng0.init_params(seed=0)
        eng0.load_params(script_subnet_alphas({k: v.clone() for k, v in eng0.named_parameters().items()}, args.depth))
        eng0.set_schedule(10.0)
        warm = torch.randn(B, 3, 224, 224, device=dev)
        eng0.step(warm, torch.zeros(B, dtype=torch.int64, device=dev), update=False)       # ranks for the plan
        eng0.grads.zero_()
        t_ev = time.perf_counter()
        eng, finished, executed = eng0.prune_event(0.2)
        torch.cuda.synchronize()
        prune_event_ms = (time.perf_counter() - t_ev) * 1e3
        assert finished and executed and (eng.Dv, eng.heads, eng.hdims, eng.hids) == (
            FT_SUBNET["embed_dim"], FT_SUBNET["heads"], FT_SUBNET["head_dims"], FT_SUBNET["hiddens"])
        del eng0, warm
        torch.cuda.empty_cache()
        eng.enter_post_search()
        import numpy as np
        ctx["mixup_fn"] = Mixup(rng=np.random.RandomState(1 + rank))
        ctx["soft"] = torch.empty(B, 1000, device=dev)
        ctx["mixed"] = torch.empty(B, 3, 224, 224, device=dev)
        step_gflop = subnet_step_gflop()
        workload = (f"post-search step (search.py:641-656: Mixup/CutMix + soft-target CE, PMIM off, decoder frozen) of the DeiT-S "
                    f"search engine after the finalising prune event (subnet of bench.py FT_SUBNET, embed {FT_SUBNET['embed_dim']}), "
                    f"depth 12, batch {B}/GPU, 224px (SURVEY 8f-3, extra configuration); prune_event() itself took "
                    f"{prune_event_ms:.0f} ms host wall")
    else:
        eng = SearchStepEngine(D, H, args.depth, B, drop_path_rate=0.1, lr=lr, device=dev, process_group=pg)
        eng.init_params(seed=0)                     # same initial weights on every rank (DDP broadcast equivalent)
        eng.set_schedule(0.0)
        step_gflop = STEP_GFLOP.get(args.model)
        workload = (f"DeiT-{args.model} bi-mask search + PMIM step, depth {args.depth}, batch {B}/GPU, 224px "
                    + ("(BASELINE.json configs[1])" if args.model == "small" else "(BASELINE.json parity / extra configuration)"))

    inner = eng.step if args.no_graph else eng.step_graphed
    eager = eng.step
    if args.workload == "post":
        def step(images, labels):          # engine.py:98-99: `samples, targets = mixup_fn(samples, targets)`, then the step
            ctx["mixup_fn"](images, labels, ctx["soft"], images_out=ctx["mixed"])
            return inner(ctx["mixed"], None, target=ctx["soft"])

        def eager_step(images, labels):
            ctx["mixup_fn"](images, labels, ctx["soft"], images_out=ctx["mixed"])
            return eager(ctx["mixed"], None, target=ctx["soft"], update=False)
    else:
        step = inner

        def eager_step(images, labels):
            return eager(images, labels, update=False)
    return eng, step, eager_step, step_gflop, workload


def extra_config(model, batch, workload, dev, steps=5, warmup=3):
    """Short device-timed run of another BASELINE.json configuration (same engine, same timing rules, fewer steps)."""
    import torch
    ns = argparse.Namespace(model=model, batch=batch, depth=12, workload=workload, no_graph=False)
    try:
        eng, step, _, step_gflop, desc = build_workload(ns, dev, None, 1, 0)
        g = torch.Generator(device="cpu").manual_seed(7)
        img = [torch.randn(batch, 3, 224, 224, generator=g).to(dev) for _ in range(2)]
        lab = [torch.randint(0, 1000, (batch,), generator=g).to(dev) for _ in range(2)]
        for i in range(warmup):
            step(img[i % 2], lab[i % 2])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(img[i % 2], lab[i % 2])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out = {"workload": desc, "value": batch / (ms / 1e3), "unit": "images/s", "ms_per_step": ms, "steps": steps,
               "warmup": warmup, "step_tflops": (step_gflop or 0) * batch / ms}
        eng.release_graphs()
        del eng, step, img, lab
    except Exception as e:                      # noqa: BLE001
        out = {"workload": f"DeiT-{model} {workload} batch {batch}", "error": f"{type(e).__name__}: {str(e)[:200]}"}
    torch.cuda.empty_cache()
    return out


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    # stray writes to fd 1 (NCCL prints its version banner there) must not end up next to the JSON line: everything but
    # the result goes to stderr
    out_fd = os.dup(1)
    os.dup2(2, 1)
    trace = os.environ.get("OFB_BENCH_TRACE")
    if trace:          # debugging aid: stage markers on stderr and a Python stack dump if a stage hangs
        import faulthandler
        faulthandler.dump_traceback_later(int(trace), repeat=False, exit=False)

    def mark(what):
        if trace:
            print(f"[bench rank {os.environ.get('RANK', '0')}] {what}", file=sys.stderr, flush=True)

    import torch
    import torch.distributed as dist
    import ofb_b200  # noqa: F401
    from ofb_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        pg = dist.group.WORLD
    dev = torch.device("cuda", local)
    B = args.batch
    eff = B * world
    eng, step, eager_step, step_gflop, workload = build_workload(args, dev, pg, world, rank)

    g = torch.Generator(device="cpu").manual_seed(1 + rank)
    n_host = 2
    host_img = [torch.randn(B, 3, 224, 224, generator=g).pin_memory() for _ in range(n_host)]
    host_lab = [torch.randint(0, 1000, (B,), generator=g).pin_memory() for _ in range(n_host)]
    dev_img = [h.to(dev) for h in host_img]
    dev_lab = [h.to(dev) for h in host_lab]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- algorithmic work of one step (GEMM / attention FLOPs, LayerNorm bytes): one logged eager pass ----------------
    ops.WORK_LOG = []
    eager_step(dev_img[0], dev_lab[0])
    work_log, ops.WORK_LOG = ops.WORK_LOG, None
    eng.grads.zero_()
    work, gemm_shapes = {}, {}
    for fam, fl, nb, dims in work_log:
        w = work.setdefault(fam, [0.0, 0.0, 0])
        w[0] += fl; w[1] += nb; w[2] += 1
        if fam == "gemm" and dims is not None:
            gemm_shapes[dims[:5]] = gemm_shapes.get(dims[:5], 0) + 1

    # ---------------- device-resident measurement ----------------
    mark("engine built")
    for i in range(args.warmup):
        step(dev_img[i % n_host], dev_lab[i % n_host])
        if trace:
            torch.cuda.synchronize()
            mark(f"warmup step {i} done")
    sync_all()
    mark("warmup done")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(dev_img[i % n_host], dev_lab[i % n_host])
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    mark("timed region done")
    launches = ops.LAUNCHES - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = eff * args.steps / (ms / 1e3)
    scal = eng.scal.cpu().tolist()

    # ---------------- N > 1: the gradient exchange in isolation (it is not overlapped by default, so this is what it adds to a step) ----------------
    exchange = None
    if world > 1 and hasattr(eng, "allreduce_grads"):
        for _ in range(3):
            eng.allreduce_grads()
        sync_all()
        n_ex = 20
        e0.record()
        for _ in range(n_ex):
            eng.allreduce_grads()
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1) / n_ex], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nbytes = eng.grads.numel() * 4
        mode = "captured in the step graph" if getattr(eng, "dp_in_graph", False) else (
            "overlapped with backward" if getattr(eng, "dp_overlap", False) else "after the graph replay, before AdamW")
        exchange = {"ms_per_step": float(t.item()), "bytes_per_rank": nbytes, "buckets": len(getattr(eng, "_dp_bounds", []) or []),
                    "gbs_per_rank": nbytes / (float(t.item()) * 1e-3) / 1e9,
                    "mode": mode, "note": "NCCL all-reduce (AVG) of the flat fp32 gradient arena, timed alone with CUDA events, max over ranks"}
        eng.grads.zero_()
    # ---------------- per-kernel durations of the same replayed steps (CUPTI) ----------------
    n_prof = min(4, args.steps)
    def prof_steps(n):
        for i in range(n):
            step(dev_img[i % n_host], dev_lab[i % n_host])
    if rank == 0:
        psampler = ClockSampler(local)          # the profiled replays are short: they may run at a higher clock than the timed region
        psampler.start()
        fam_ms, kern = kernel_breakdown(prof_steps, n_prof)
        clocks_profiled = psampler.stop()
    else:
        prof_steps(n_prof)          # the steps of an N > 1 run hold the gradient exchange: every rank runs them, rank 0 records
        fam_ms, kern = None, "skipped"
    sync_all()
    mark("roofline pass done")
    # ---------------- end to end: host buffers -> step -> loss on host ----------------
    copy_stream = torch.cuda.Stream(device=dev)
    stage_img = [torch.empty_like(dev_img[0]) for _ in range(2)]
    stage_lab = [torch.empty_like(dev_lab[0]) for _ in range(2)]
    loss_host = torch.zeros(8).pin_memory()
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    # diagnostic only (never set for a reported number): "nocopy" skips the image H2D copy, "unused" copies but steps on the
    # resident batch - separates copy/compute contention from stream dependencies when e2e lags the device-timed value
    e2e_variant = os.environ.get("OFB_E2E_VARIANT", "")

    def prefetch(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s])
            if e2e_variant != "nocopy":
                stage_img[s].copy_(host_img[i % n_host], non_blocking=True)
            stage_lab[s].copy_(host_lab[i % n_host], non_blocking=True)
            ready[s].record(copy_stream)

    def e2e_loop(n):
        for s in range(2):
            consumed[s].record()
        prefetch(0)
        for i in range(n):
            if i + 1 < n:
                prefetch(i + 1)
            s = i % 2
            torch.cuda.current_stream().wait_event(ready[s])
            step(dev_img[s] if e2e_variant == "unused" else stage_img[s], stage_lab[s])
            consumed[s].record()
            loss_host.copy_(eng.scal, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_loop(max(2, min(args.warmup, 3)))
    sync_all()
    mark("e2e warmup done")
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = eff * args.steps / (float(t.item()) / 1e3)
    h2d = host_img[0].numel() * 4 + host_lab[0].numel() * 8

    if rank == 0:
        pk = peaks()
        tr = gemm_traffic()
        step_ms = ms / args.steps
        timed_over = f"{n_prof} CUDA-graph replays of the timed step, CUPTI kernel records (torch.profiler), grouped by kernel family"
        if fam_ms is None:
            timed_over = f"unavailable ({kern})"
            fam_ms, kern = {}, {}

        def fam(name):
            return fam_ms.get(name, 0.0)

        def frac(x, peak):
            return x / peak if x is not None else None
        gemm_tf = work.get("gemm", [0, 0, 0])[0] / (fam("gemm") / 1e3) / 1e12 if fam("gemm") > 0 else None
        attn = {}
        for k in ("attn_fwd", "attn_bwd"):
            fl, nb, n = work.get(k, [0.0, 0.0, 0])
            if fam(k) > 0:
                attn[k] = {"us_per_launch": fam(k) * 1e3 / max(n, 1), "launches_per_step": n,
                           "tflops": fl / (fam(k) / 1e3) / 1e12, "frac_tensor": fl / (fam(k) / 1e3) / 1e12 / pk["tflops"],
                           "gbs": nb / (fam(k) / 1e3) / 1e9, "frac_hbm": nb / (fam(k) / 1e3) / 1e9 / pk["hbm"],
                           "share_of_step": fam(k) / step_ms}
        ln_b = work.get("ln", [0, 0, 0])
        ln_gbs = ln_b[1] / (fam("ln") / 1e3) / 1e9 if fam("ln") > 0 else None
        step_tf = (step_gflop or 0) * value / 1e3 / world
        ksum = sum(fam_ms.values())
        line = {
            "metric": METRIC[args.workload],
            "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload, "global_batch": eff, "parallelism": f"dp{world}",
                       "l2": "activations per step (>7 GB) exceed the 126 MB L2; no explicit flush",
                       "launch": "host launches" if args.no_graph else "CUDA graph replay (one graph per input buffer)",
                       "step_gflop_per_image": step_gflop},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 32},
            "gpu_launches": launches,
            # dominant kernel family: the tcgen05 GEMM (all fused epilogues). achieved = algorithmic FLOPs of a step's GEMM
            # launches (2 M N K, unpadded) / their summed device durations in the replayed graph
            "roofline": {"bound": "tensor", "kernel": "ofb::gemm_kernel (tcgen05, all epilogues)", "achieved": gemm_tf,
                         "peak": pk["tflops"], "unit": "TFLOP/s", "frac": frac(gemm_tf, pk["tflops"]),
                         "traffic": (tr["dram_bytes_per_launch"] if tr and args.workload == "search" and args.model == "small"
                                     and B == 256 and args.depth == 12 else None),
                         "traffic_note": "DRAM bytes per GEMM launch (mean over the launches of a step), ncu --set full, "
                                         "profiles/r02c_gemm_traffic.json",
                         "peak_source": pk["src"], "launches_per_step": work.get("gemm", [0, 0, 0])[2],
                         "ms_per_step": fam("gemm"), "share_of_step": fam("gemm") / step_ms if step_ms > 0 else None,
                         "share_of_kernel_time": fam("gemm") / ksum if ksum > 0 else None,
                         "timed_over": timed_over, "clocks_profiled": clocks_profiled},
            # the whole step against the tensor roofline (north_star: >= 0.60): algorithmic FLOPs of the step / step time
            "roofline_step": {"bound": "tensor", "achieved": step_tf, "peak": pk["tflops"], "unit": "TFLOP/s",
                              "frac": step_tf / pk["tflops"], "target_frac": 0.60},
            # attention kernels: tensor work AND algorithmic HBM bytes (q, k, v, o [, dO, dq, dk, dv]); the smaller time floor is HBM
            "roofline_attention": attn,
            # the HBM-bound kernel family of the step: LayerNorm forward / backward, algorithmic bytes (bf16 rows
            # read + written) over device durations, against the measured copy bandwidth
            "roofline_hbm": {"bound": "hbm", "kernel": "ofb::ln_fwd*/ln_bwd* (LayerNorm forward + backward)",
                             "achieved": ln_gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": frac(ln_gbs, pk["hbm"]),
                             "launches_per_step": ln_b[2], "ms_per_step": fam("ln"),
                             "share_of_step": fam("ln") / step_ms if step_ms > 0 else None, "traffic": None},
            # every kernel family of the replayed step: ms per step and share of the device-timed step (branches of the graph
            # overlap, so the shares can add up to slightly more than 1)
            "kernel_shares": {k: {"ms_per_step": v, "share_of_step": v / step_ms} for k, v in sorted(fam_ms.items(),
                                                                                                      key=lambda kv: -kv[1])},
            "kernel_sum_ms": ksum,
            "top_kernels": [{"name": k[:96], "ms_per_step": v[0], "launches_per_step": v[1]}
                            for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0])[:12]],
            "losses": {"base": scal[0], "arch": scal[1], "decoder": scal[2], "total": scal[3]},
        }
        if exchange is not None:
            line["exchange"] = exchange
        if world == 1:
            eng.release_graphs()
            del eng, step, eager_step, stage_img, dev_img
            torch.cuda.empty_cache()
            cb_ms, cb_err = cublas_same_shapes(gemm_shapes, dev)
            line["roofline"]["cublas_same_shapes"] = {
                "ms_per_step": cb_ms, "ours_ms_per_step": fam("gemm"), "ours_over_cublas_time": (fam("gemm") / cb_ms) if cb_ms else None,
                "note": "torch.matmul bf16 on the step's GEMM shapes and operand layouts, each timed in isolation, plain store "
                        "(no fused gate / GELU / residual / d-gate / split-K epilogue); our figure includes those epilogues",
                "error": cb_err}
        if world == 1 and not args.no_cpu_baseline:
            rate, cores, dt, kind = cpu_arm(args)
            what = "unmodified reference, oracle/_ref" if kind == "reference" else "oracle port"
            line["cpu_baseline"] = {"value": rate, "unit": "images/s", "cores": cores, "kind": kind,
                                    "sample": f"{args.cpu_sample_batch} images / step of the same DeiT-{args.model} {args.workload} "
                                              f"step ({what}, fp32, {cores} threads, {dt:.1f} s/step)"}
        if world == 1 and not args.no_eager and args.workload == "search":
            line["gpu_eager_baseline"] = gpu_eager_baseline(args.model, args.depth, B)
        if world == 1 and not args.no_extra and args.workload == "search" and args.model == "small":
            line["extra_configs"] = [extra_config("tiny", 1024, "search", dev), extra_config("base", 256, "search", dev),
                                     extra_config("small", 256, "finetune", dev)]
        sys.stdout.flush()
        os.write(out_fd, (json.dumps(line) + "\n").encode())
    mark("result written")
    if world > 1:
        # graphs that captured NCCL launches must be gone before the communicator is torn down (destroy_process_group
        # otherwise blocks); the timer guarantees the process ends even if the teardown stalls
        eng.release_graphs()
        torch.cuda.synchronize()
        threading.Timer(30.0, os._exit, (0,)).start()
        dist.destroy_process_group()
        os._exit(0)


if __name__ == "__main__":
    main()
