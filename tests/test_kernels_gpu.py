"""GPU parity tests of the individual sm_100a kernels, called through the C ABI (ofb_b200.ops), against plain fp32
PyTorch restatements / the CPU oracle of the same reference operation.

Tolerances: bf16 paths rel 2e-2 (north_star), fp32 paths rel 1e-4, index / mask outputs exact."""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16_TOL = 2e-2
FP32_TOL = 1e-4


def rel(got, ref):
    got, ref = got.float(), ref.float()
    return float((got - ref).abs().max() / (ref.abs().max() + 1e-12))


def rnd(*shape, s=1.0, dev="cuda"):
    return (torch.randn(*shape, device=dev) * s).to(torch.bfloat16)


# ---------------------------------------------------------------------------------------------------------------------
# GEMM epilogues (nn.Linear fwd / dgrad / wgrad of layers.py:491, 515, 845-863)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,bn", [(256, 128, 64, 128), (1576, 1152, 384, 192), (1576, 1536, 384, 256),
                                      (1576, 384, 1536, 64), (256, 1000, 384, 0), (394, 192, 192, 0)])
def test_gemm_store(cuda_dev, M, N, K, bn):
    from ofb_b200 import ops
    torch.manual_seed(0)
    A, B = rnd(M, K), rnd(N, K, s=0.05)
    bias, gate = torch.randn(N, device="cuda"), torch.rand(N, device="cuda") + 0.5
    res, rs = rnd(M, N), torch.rand((M + 196) // 197, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ref = A.float() @ B.float().t()
    rows = torch.arange(M, device="cuda") // 197
    ops.gemm(ops.EPI_STORE, A, B, M=M, N=N, K=K, out0=out, bias=bias, colscale=gate, rowscale=rs, rows_per_scale=197,
             res=res, bn=bn)
    assert rel(out, rs[rows, None] * (ref + bias) * gate + res.float()) < BF16_TOL
    ops.gemm(ops.EPI_STORE, A, B, M=M, N=N, K=K, out0=out, bias=bias, rowscale=rs, rows_per_scale=197, res=res, bn=bn,
             bias_rowscaled=True)
    assert rel(out, ref + rs[rows, None] * bias + res.float()) < BF16_TOL
    outf = torch.empty(M, N, device="cuda")
    ops.gemm(ops.EPI_STORE, A, B, M=M, N=N, K=K, out0=outf, out_fp32=True, bias=bias, bn=bn)
    assert rel(outf, ref + bias) < 1e-3   # bf16 inputs, fp32 accumulate/out
    # MN-major B (data-gradient form: B given as [K, N])
    Bt = B.t().contiguous()
    ops.gemm(ops.EPI_STORE, A, Bt, M=M, N=N, K=K, out0=outf, out_fp32=True, b_mn=True, bn=bn)
    assert rel(outf, ref) < 1e-3


def test_gemm_strided_rows(cuda_dev):
    """head forward / backward read and write the cls rows in place (row stride T*D)."""
    from ofb_b200 import ops
    torch.manual_seed(1)
    Bsz, T, D, Cn = 8, 197, 384, 1000
    lat = rnd(Bsz * T, D)
    W = rnd(Cn, D, s=0.05)
    logits = torch.empty(Bsz, Cn, device="cuda")
    ops.gemm(ops.EPI_STORE, lat, W, M=Bsz, N=Cn, K=D, out0=logits, out_fp32=True, lda=T * D)
    assert rel(logits, lat.view(Bsz, T, D)[:, 0].float() @ W.float().t()) < 1e-3
    dl = rnd(Bsz, Cn, s=0.01)
    dlat = torch.zeros(Bsz * T, D, device="cuda", dtype=torch.bfloat16)
    ops.gemm(ops.EPI_STORE, dl, W, M=Bsz, N=D, K=Cn, out0=dlat, ld0=T * D, b_mn=True)
    got = dlat.view(Bsz, T, D)
    assert rel(got[:, 0], dl.float() @ W.float()) < BF16_TOL
    assert got[:, 1:].abs().max() == 0


def test_gemm_fc1_and_fc2_dgrad(cuda_dev):
    """bi-masked MLP (layers.py:845-863) with the hidden activations kept transposed [hidden, tokens]."""
    from ofb_b200 import ops
    torch.manual_seed(2)
    M, D, Hd = 1571, 384, 1536            # ragged token count: the token pitch is padded to 8
    ldT = (M + 7) // 8 * 8
    x, W1 = rnd(M, D), rnd(Hd, D, s=0.05)
    b1, gate = torch.randn(Hd, device="cuda") * 0.1, torch.rand(Hd, device="cuda") + 0.3
    rs = (torch.rand((M + 196) // 197, device="cuda") > 0.3).float() / 0.7
    rows = torch.arange(M, device="cuda") // 197
    u = torch.zeros(Hd, ldT, device="cuda", dtype=torch.bfloat16)
    h = torch.zeros_like(u)
    ops.gemm(ops.EPI_FC1, W1, x, M=Hd, N=M, K=D, out0=u, out1=h, bias=b1, colscale=gate, rowscale=rs, rows_per_scale=197, bn=256)
    ru = x.float() @ W1.float().t() + b1
    assert rel(u[:, :M].t(), ru) < BF16_TOL
    assert rel(h[:, :M].t(), rs[rows, None] * F.gelu(ru * gate)) < BF16_TOL
    # (the pad columns M..ldT may be written: TMA clips stores at 16-byte granularity; every consumer bounds tokens by M)
    # fc2 forward consumes h^T as an MN-major A operand
    W2 = rnd(D, Hd, s=0.05)
    res = rnd(M, D)
    y = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    ops.gemm(ops.EPI_STORE, h, W2, M=M, N=D, K=Hd, out0=y, res=res, a_mn=True)
    assert rel(y, h[:, :M].t().float() @ W2.float().t() + res.float()) < BF16_TOL
    # backward of  y = rs * gelu(u*g) @ W2^T :  dh = rs * (dy @ W2)
    dy = rnd(M, D)
    du = torch.zeros_like(u)
    R = ops.gemm_mlp_partial_rows(M, 256)
    p0, p1 = torch.zeros(R, Hd, device="cuda"), torch.zeros(R, Hd, device="cuda")
    ops.gemm(ops.EPI_FC2_DGRAD, W2, dy, M=Hd, N=M, K=D, out0=du, aux=u, colscale=gate, rowscale=rs, rows_per_scale=197,
             colpart0=p0, colpart1=p1, a_mn=True, bn=256)
    uf = u[:, :M].t().float().requires_grad_(True)
    gf = gate.clone().requires_grad_(True)
    F.gelu(uf * gf).backward(rs[rows, None] * (dy.float() @ W2.float()))
    assert rel(du[:, :M].t(), uf.grad) < BF16_TOL
    assert rel(p0.sum(0), gf.grad) < 1e-3
    assert rel(p1.sum(0), uf.grad.sum(0)) < 1e-3
    # weight gradients / data gradient with the transposed operands
    dW2 = torch.zeros(D, Hd, device="cuda")
    ops.gemm(ops.EPI_WGRAD, dy, h, M=D, N=Hd, K=M, out0=dW2, a_mn=True)
    assert rel(dW2, dy.float().t() @ h[:, :M].t().float()) < 1e-3
    dW1 = torch.zeros(Hd, D, device="cuda")
    ops.gemm(ops.EPI_WGRAD, du, x, M=Hd, N=D, K=M, out0=dW1, b_mn=True)
    assert rel(dW1, du[:, :M].float() @ x.float()) < 1e-3
    dx = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    ops.gemm(ops.EPI_STORE, du, W1, M=M, N=D, K=Hd, out0=dx, a_mn=True, b_mn=True, res=res)
    assert rel(dx, du[:, :M].t().float() @ W1.float() + res.float()) < BF16_TOL


@pytest.mark.parametrize("R,NO,KI", [(128, 128, 128), (1576, 1536, 384), (1576, 1000, 384), (1576, 384, 768)])
def test_gemm_wgrad(cuda_dev, R, NO, KI):
    from ofb_b200 import ops
    torch.manual_seed(3)
    dY, X = rnd(R, NO), rnd(R, KI)
    dW = torch.ones(NO, KI, device="cuda")
    sc = torch.tensor([0.5], device="cuda")
    ops.gemm(ops.EPI_WGRAD, dY, X, M=NO, N=KI, K=R, out0=dW, a_mn=True, b_mn=True, scale_ptr=sc)
    assert rel(dW, 1 + 0.5 * (dY.float().t() @ X.float())) < 1e-3


@pytest.mark.parametrize("R,N,ld", [(50432, 768, 768), (1571, 384, 392), (256, 1000, 1000), (197, 250, 250)])
def test_colsum_bf16(cuda_dev, R, N, ld):
    """bias gradients as column sums of a bf16 matrix (decoder / head bias: engine.py:169 autograd of vt:723, 744): the
    16-byte-load kernel (N % 8 == 0) and the 4-byte fallback, accumulating into `out`, bit-identical from run to run."""
    from ofb_b200 import ops
    torch.manual_seed(11)
    x = torch.zeros(R, ld, device="cuda", dtype=torch.bfloat16)
    x[:, :N] = rnd(R, N)
    sc = torch.tensor([0.5], device="cuda")
    out = torch.ones(N, device="cuda")
    ops.colsum_bf16(x, R, N, out, scale=2.0, scale_dev=sc, ld=ld)
    ref = 1 + x[:, :N].float().sum(0)
    assert rel(out, ref) < 1e-5
    out2 = torch.ones(N, device="cuda")
    ops.colsum_bf16(x, R, N, out2, scale=2.0, scale_dev=sc, ld=ld)
    assert torch.equal(out, out2)


# ---------------------------------------------------------------------------------------------------------------------
# LayerNorm (layers.py:96-98)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,D", [(1576, 384), (788, 192), (394, 768), (5, 96)])
def test_layernorm_fwd_bwd(cuda_dev, M, D):
    from ofb_b200 import ops
    torch.manual_seed(4)
    x = rnd(M, D, s=2.0)
    gamma, beta = 1 + 0.1 * torch.randn(D, device="cuda"), 0.1 * torch.randn(D, device="cuda")
    y = torch.empty_like(x)
    mean, rstd = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
    ops.layernorm_fwd(x, gamma, beta, y, mean, rstd, 1e-6)
    xf = x.float().requires_grad_(True)
    gf, bf = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.layer_norm(xf, (D,), gf, bf, 1e-6)
    assert rel(y, ref) < BF16_TOL
    assert rel(mean, xf.mean(-1)) < 1e-4
    dy = rnd(M, D)
    ref.backward(dy.float())
    R = ops.layernorm_bwd_parts(M)
    dx = torch.empty_like(x)
    pg, pb, pd = (torch.zeros(R, D, device="cuda") for _ in range(3))
    rs = torch.rand((M + 196) // 197, device="cuda")
    ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, pg, pb, pd, rs, 197)
    assert rel(dx, xf.grad) < BF16_TOL
    assert rel(pg.sum(0), gf.grad) < 1e-3
    assert rel(pb.sum(0), bf.grad) < 1e-3
    rows = torch.arange(M, device="cuda") // 197
    assert rel(pd.sum(0), (rs[rows, None] * dx.float()).sum(0)) < 1e-2
    out = torch.ones(D, device="cuda")
    ops.reduce_partials(pg, R, D, out, scale=2.0, div_by=gamma, accumulate=True)
    assert rel(out, 1 + 2.0 * pg.sum(0) / gamma) < 1e-5


@pytest.mark.parametrize("M,D,Dv,with_res", [(1576, 288, 288, True), (788, 256, 252, True), (1576, 208, 204, False),
                                              (394, 336, 336, False), (1000, 104, 102, True), (394, 376, 372, True),
                                              (197, 72, 66, False), (394, 576, 576, True), (394, 200, 198, True)])
def test_layernorm_pruned_widths(cuda_dev, M, D, Dv, with_res):
    """LayerNorm over the Dv real channels of a zero-padded pruned embedding (physical width D = Dv rounded up to 8), forward and
    backward incl. the residual-branch gradient of pre-norm blocks: the word-granular packed kernels (any even Dv)."""
    from ofb_b200 import ops
    torch.manual_seed(D + Dv)
    x = rnd(M, D, s=2.0)
    x[:, Dv:] = 0
    gamma, beta = 1 + 0.1 * torch.randn(D, device="cuda"), 0.1 * torch.randn(D, device="cuda")
    gamma[Dv:] = 0
    beta[Dv:] = 0
    y = torch.full_like(x, 7.0)
    mean, rstd = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
    ops.layernorm_fwd(x, gamma, beta, y, mean, rstd, 1e-6, d_valid=Dv)
    xf = x[:, :Dv].float().requires_grad_(True)
    gf, bf = gamma[:Dv].clone().requires_grad_(True), beta[:Dv].clone().requires_grad_(True)
    ref = F.layer_norm(xf, (Dv,), gf, bf, 1e-6)
    assert rel(y[:, :Dv], ref) < BF16_TOL and float(y[:, Dv:].abs().max() if Dv < D else 0) == 0.0
    assert rel(mean, xf.mean(-1)) < 1e-4
    assert rel(rstd, (xf.var(-1, unbiased=False) + 1e-6).rsqrt()) < 1e-4
    dy = rnd(M, D)
    dy[:, Dv:] = 0
    ref.backward(dy[:, :Dv].float())
    dres = rnd(M, D) if with_res else None
    if dres is not None:
        dres[:, Dv:] = 0
    R = ops.layernorm_bwd_parts(M)
    dx = torch.full_like(x, 7.0)
    pg, pb, pd = (torch.zeros(R, D, device="cuda") for _ in range(3))
    rs = torch.rand((M + 196) // 197, device="cuda")
    ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, pg, pb, pd, rs, 197, dres=dres, d_valid=Dv)
    want = xf.grad + (dres[:, :Dv].float() if dres is not None else 0)
    assert rel(dx[:, :Dv], want) < BF16_TOL
    assert float(dx[:, Dv:].abs().max() if Dv < D else 0) == 0.0            # padding stays exactly zero
    assert rel(pg.sum(0)[:Dv], gf.grad) < 1e-3 and rel(pb.sum(0)[:Dv], bf.grad) < 1e-3
    assert float(pg.sum(0)[Dv:].abs().max() if Dv < D else 0) == 0.0
    rows = torch.arange(M, device="cuda") // 197
    assert rel(pd.sum(0), (rs[rows, None] * dx.float()).sum(0)) < 1e-2


# ---------------------------------------------------------------------------------------------------------------------
# attention (layers.py:507-514) forward / backward incl. the bi-mask gate products
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,T", [(2, 3, 197), (3, 6, 197), (1, 2, 64)])
def test_attention_fwd_bwd(cuda_dev, B, H, T):
    from ofb_b200 import ops
    torch.manual_seed(5)
    d, D = 64, H * 64
    scale = d ** -0.5
    gate = torch.rand(D, device="cuda") * 0.5 + 0.5
    qkv = rnd(B, T, 3, H, d, s=1.0)
    ds = (torch.rand(B, device="cuda") > 0.3).float() / 0.7
    ds[0] = 1 / 0.7
    o = torch.empty(B, T, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, T, device="cuda")
    ops.attention_fwd(qkv, o, lse, ds, B, T, H, scale)
    qf = qkv.float().requires_grad_(True)
    q, k, v = qf.permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-2, -1)) * scale
    oref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, T, D) * ds.view(B, 1, 1)
    assert rel(o, oref) < BF16_TOL
    assert rel(lse, torch.logsumexp(s, -1)) < 1e-3
    d_o = rnd(B, T, D, s=0.5)
    oref.backward(d_o.float())
    dqkv = torch.empty_like(qkv)
    pg = torch.zeros(B, D, device="cuda")
    pb = torch.zeros(B, 3 * D, device="cuda")
    # contract: d_o is the gradient w.r.t. the un-scaled attention output (the proj data-gradient GEMM applies DropPath)
    d_o_in = (d_o.float() * ds.view(B, 1, 1)).to(torch.bfloat16)
    ops.attention_bwd(qkv, o, d_o_in, lse, gate, ds, dqkv, pg, pb, B, T, H, scale)
    g = qf.grad                                             # grad w.r.t. the gated q,k,v  [B,T,3,H,d]
    gview = gate.view(1, 1, 1, H, d)
    assert rel(dqkv, g * gview) < BF16_TOL                  # d pre-gate
    assert rel(pg.sum(0), (g * qf.detach()).sum((0, 1, 2)).reshape(-1)) < BF16_TOL
    assert rel(pb.sum(0), (g * gview).sum((0, 1)).reshape(-1)) < BF16_TOL


@pytest.mark.parametrize("amp,shift", [(1.0, 0.0), (6.0, 0.0), (1.0, 14.0), (1.0, -14.0)],
                         ids=["normal", "wide_logits", "logits_plus_200", "logits_minus_200"])
def test_attention_fwd_single_pass_guard(cuda_dev, amp, shift):
    """The forward softmax runs in one pass without subtracting the row maximum; rows whose sum leaves the safe exponent range
    fall back to the two-pass path. Logits of ordinary size, a +-150 spread, and rows pushed to +-200 (where exp over / under
    flows without a reference point) must all match torch's softmax."""
    from ofb_b200 import ops
    torch.manual_seed(11)
    B, H, T, d = 2, 3, 197, 64
    D = H * d
    qkv = torch.randn(B, T, 3, H, d, device="cuda") * amp
    # constant component along one channel: q . k gets an additive shift of 64 * shift^2 / 8 in half of the rows
    qkv[:, ::2, 0, :, 0] += shift * 8.0
    qkv[:, :, 1, :, 0] += abs(shift) * 8.0 if shift else 0.0
    qkv = qkv.to(torch.bfloat16)
    o = torch.empty(B, T, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, T, device="cuda")
    ops.attention_fwd(qkv, o, lse, None, B, T, H, d ** -0.5)
    q, k, v = qkv.float().permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-2, -1)) * d ** -0.5
    if shift:
        assert float(s.abs().max()) > 150
    oref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, T, D)
    assert torch.isfinite(o.float()).all() and torch.isfinite(lse).all()
    assert rel(o, oref) < BF16_TOL
    assert float((lse - torch.logsumexp(s, -1)).abs().max()) < 1e-3 * max(1.0, float(s.abs().max()))


# ---------------------------------------------------------------------------------------------------------------------
# token assembly, PMIM, targets
# ---------------------------------------------------------------------------------------------------------------------
def test_patchify_mask_droppath_cls(cuda_dev):
    from ofb_b200 import ops
    from ofb_oracle import pmim_mask
    torch.manual_seed(6)
    B, D, L = 3, 192, 196
    img = torch.randn(B, 3, 224, 224, device="cuda")
    pat = torch.empty(B * L, 768, device="cuda", dtype=torch.bfloat16)
    ops.patchify(img, pat)
    ref = img.reshape(B, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(B * L, 768)
    assert torch.equal(pat, ref.to(torch.bfloat16))
    noise = torch.rand(B, L, device="cuda")
    noise[0, 5] = noise[0, 9]    # a tie
    mask = torch.empty(B, L, device="cuda")
    for keep in (186, 166, 147):
        ops.pmim_mask(noise, mask, keep)
        assert torch.equal(mask.cpu(), pmim_mask(noise.cpu(), keep))
    u = torch.rand(24, B, device="cuda")
    p = torch.linspace(0, 0.1, 12, device="cuda").repeat_interleave(2)
    sc = torch.empty(24, B, device="cuda")
    ops.droppath_scale(u, p, sc)
    keep = (1 - p).view(-1, 1)
    assert torch.allclose(sc, torch.floor(keep + u) / keep)
    cls, pos, gate = torch.randn(D, device="cuda"), torch.randn(197, D, device="cuda"), torch.rand(D, device="cuda")
    x = torch.zeros(B, 197, D, device="cuda", dtype=torch.bfloat16)
    ops.cls_rows(cls, pos, gate, x, B, 197, D)
    assert rel(x[:, 0], ((cls + pos[0]) * gate).expand(B, D)) < BF16_TOL and x[:, 1:].abs().max() == 0


def test_norm_targets(cuda_dev):
    from ofb_b200 import ops
    from ofb_oracle import norm_targets, patchify_pixel_shuffle
    torch.manual_seed(7)
    B, L = 2, 196
    img = torch.randn(B, 3, 224, 224) + 2 * F.interpolate(torch.randn(B, 3, 7, 7), size=224, mode="bilinear")
    mask = (torch.rand(B, L) < 0.3).float()
    tgt = torch.zeros(B * L, 768, device="cuda")
    ops.norm_targets(img.cuda(), mask.cuda(), tgt)
    ref = patchify_pixel_shuffle(norm_targets(img, 47)).reshape(B * L, 768)
    m = mask.reshape(-1).bool()
    assert (tgt.cpu()[m] - ref[m]).abs().max() < 2e-4       # fp32, |target| ~ O(1..5)
    assert tgt.cpu()[~m].abs().max() == 0                   # unmasked patches are never needed / never written


@pytest.mark.parametrize("B,HW,frac", [(3, 48, 1.0), (5, 32, 0.5), (2, 64, 0.0), (1, 224, 1.0)])
def test_norm_targets_edges(cuda_dev, B, HW, frac):
    """Images smaller than the 47-pixel window (every window is clipped by the border on all sides: vt:121-141 divides by the
    clipped pixel count), patch counts that are not a multiple of the 4 patches a CTA takes, everything / nothing masked."""
    from ofb_b200 import ops
    from ofb_oracle import norm_targets, patchify_pixel_shuffle
    torch.manual_seed(23)
    L = (HW // 16) ** 2
    img = torch.randn(B, 3, HW, HW) * 1.5 + 0.3
    mask = (torch.rand(B, L) < frac).float()
    tgt = torch.zeros(B * L, 768, device="cuda")
    ops.norm_targets(img.cuda(), mask.cuda(), tgt)
    ref = patchify_pixel_shuffle(norm_targets(img, 47)).reshape(B * L, 768)
    m = mask.reshape(-1).bool()
    if m.any():
        assert (tgt.cpu()[m] - ref[m]).abs().max() < 2e-4
    if (~m).any():
        assert tgt.cpu()[~m].abs().max() == 0


@pytest.mark.parametrize("n_dec,n_mask", [(40, 37 * 196), (9457, 27), (1, 3), (4099, 50176)])
def test_loss_finalize_vector_lengths(cuda_dev, n_dec, n_mask):
    """The two long sums of loss_finalize (decoder partials, PMIM mask) at lengths around the 16-byte / four-in-flight chunks."""
    from ofb_b200 import ops
    torch.manual_seed(29)
    rows = torch.rand(8, device="cuda") + 1.0
    dec_part = torch.rand(n_dec, device="cuda")
    mask = (torch.rand(n_mask, device="cuda") < 0.4).float()
    mask[0] = 1.0
    arch = torch.tensor([0.75], device="cuda")
    scal = torch.zeros(8, device="cuda")
    ops.loss_finalize(rows, dec_part, mask, arch, 1.0, scal)
    base = rows.double().mean().item()
    msum = mask.double().sum().item()
    denom = (msum * 256 + 1e-5) * 3
    dec = dec_part.double().sum().item() / denom
    want = [base, 0.75, dec, base + 0.75 + base, base / dec, base / dec / denom, msum]
    assert rel(scal[:7], torch.tensor(want, device="cuda", dtype=torch.float32)) < 1e-4


@pytest.mark.parametrize("B,D", [(3, 192), (5, 288), (2, 384), (17, 768), (3, 200)])
def test_patch_embed_epilogue_and_embed_bwd(cuda_dev, B, D):
    """D = 192 / 288 / 384 / 768: embed_bwd splits the embedding columns over three CTAs per token position (8 / 12 / 16 / 32
    16-byte vectors each); D = 200 (25 vectors) has no admissible split and runs one CTA per position."""
    from ofb_b200 import ops
    torch.manual_seed(8)
    L, T = 196, 197
    pat, W = rnd(B * L, 768), rnd(D, 768, s=0.05)
    bias, gate = torch.randn(D, device="cuda") * 0.1, torch.rand(D, device="cuda") * 0.5 + 0.5
    pos, mt = torch.randn(T, D, device="cuda") * 0.1, torch.randn(D, device="cuda") * 0.1
    mask = (torch.rand(B * L, device="cuda") < 0.2).float()
    x0 = torch.zeros(B * T, D, device="cuda", dtype=torch.bfloat16)
    ops.gemm(ops.EPI_PATCH, pat, W, M=B * L, N=D, K=768, out0=x0, bias=bias, colscale=gate, pos=pos, mask_token=mt,
             rowmask=mask, tokens=L)
    conv = (pat.float() @ W.float().t() + bias).view(B, L, D)
    m3 = mask.view(B, L, 1)
    ref = ((conv + pos[1:]) * (1 - m3) + m3 * mt) * gate
    got = x0.view(B, T, D)
    assert rel(got[:, 1:], ref) < BF16_TOL and got[:, 0].abs().max() == 0
    # backward
    g0 = rnd(B * T, D)
    x0r = rnd(B * T, D)
    dconv = torch.empty(B * L, D, device="cuda", dtype=torch.bfloat16)
    pgx, ppos, pmt = (torch.zeros(T, D, device="cuda") for _ in range(3))
    ops.embed_bwd(g0, x0r, gate, mask, dconv, pgx, ppos, pmt, B, T, D)
    g3, x3 = g0.float().view(B, T, D), x0r.float().view(B, T, D)
    assert rel(dconv.view(B, L, D), g3[:, 1:] * gate * (1 - m3)) < BF16_TOL
    assert rel(pgx, (g3 * x3).sum(0)) < 1e-4
    want_pos = torch.cat([(g3[:, :1] * gate).sum(0), (g3[:, 1:] * gate * (1 - m3)).sum(0)])
    assert rel(ppos, want_pos) < 1e-4
    assert rel(pmt[1:], (g3[:, 1:] * gate * m3).sum(0)) < 1e-4


def test_decoder_epilogue(cuda_dev):
    from ofb_b200 import ops
    torch.manual_seed(9)
    B, D, L, T = 2, 192, 196, 197
    lat, W = rnd(B * T, D), rnd(768, D, s=0.1)
    bias = torch.randn(768, device="cuda") * 0.1
    mask = (torch.rand(B * L, device="cuda") < 0.2).float()
    tgt = torch.randn(B * L, 768, device="cuda")
    sgn = torch.empty(B * T, 768, device="cuda", dtype=torch.bfloat16)
    bn = 256
    parts = torch.zeros(((B * T + 127) // 128) * (768 // bn) * 8, device="cuda")
    ops.gemm(ops.EPI_DECODER, lat, W, M=B * T, N=768, K=D, out0=sgn, bias=bias, rowmask=mask, target=tgt, tokens=L,
             colpart0=parts, bn=bn)
    rec = (lat.float() @ W.float().t() + bias).view(B, T, 768)[:, 1:]
    diff = (rec - tgt.view(B, L, 768)) * mask.view(B, L, 1)
    assert rel(parts.sum(), diff.abs().sum()) < 1e-3
    got = sgn.float().view(B, T, 768)
    assert got[:, 0].abs().max() == 0
    big = diff.abs() > 0.05                       # sign is only well defined away from 0 (bf16 inputs)
    assert torch.equal(got[:, 1:][big], torch.sign(diff)[big])
    assert got[:, 1:][mask.view(B, L) == 0].abs().max() == 0


# ---------------------------------------------------------------------------------------------------------------------
# losses and optimizer
# ---------------------------------------------------------------------------------------------------------------------
def test_cross_entropy_and_finalize(cuda_dev):
    from ofb_b200 import ops
    torch.manual_seed(10)
    B, Cn = 37, 1000
    logits = torch.randn(B, Cn, device="cuda") * 2
    labels = torch.randint(0, Cn, (B,), device="cuda")
    rows = torch.empty(B, device="cuda")
    dl = torch.empty(B, Cn, device="cuda", dtype=torch.bfloat16)
    ops.ls_cross_entropy(logits, labels, rows, dl, 0.1, 1.0)
    lf = logits.clone().requires_grad_(True)
    logp = F.log_softmax(lf, -1)
    loss = (0.9 * -logp.gather(1, labels[:, None]).squeeze(1) + 0.1 * -logp.mean(-1))
    assert rel(rows, loss) < FP32_TOL
    loss.mean().backward()
    assert rel(dl, lf.grad) < BF16_TOL
    dec_part = torch.rand(40, device="cuda")
    mask = (torch.rand(B * 196, device="cuda") < 0.1).float()
    arch = torch.tensor([3.25], device="cuda")
    scal = torch.zeros(8, device="cuda")
    ops.loss_finalize(rows, dec_part, mask, arch, 0.5, scal)
    base = loss.mean().item()
    denom = (mask.sum().item() * 256 + 1e-5) * 3
    dec = dec_part.sum().item() / denom
    want = [base, 3.25, dec, base + 3.25 + base, base / dec, base / dec / denom * 0.5, mask.sum().item()]
    assert rel(scal[:7], torch.tensor(want, device="cuda")) < 1e-4


def test_adamw_matches_oracle(cuda_dev):
    from ofb_b200 import ops
    from ofb_oracle import adamw_step
    torch.manual_seed(11)
    n = 4 * 1000 + 4 * 37
    p, g = torch.randn(n), torch.randn(n) * 0.01
    m, v = torch.randn(n) * 0.01, torch.rand(n) * 1e-4
    ends = [4 * 500, 4 * 1000, n]
    hp = [dict(lr=1e-3, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8),
          dict(lr=1e-3, weight_decay=1e-3, betas=(0.9, 0.999), eps=1e-8),
          dict(lr=2e-3, weight_decay=1e-3, betas=(0.5, 0.999), eps=1e-8)]
    step = 3
    hyper = torch.zeros(3, 8)
    for i, h in enumerate(hp):
        hyper[i, :7] = torch.tensor([h["lr"], h["weight_decay"], h["betas"][0], h["betas"][1], h["eps"],
                                     1 - h["betas"][0] ** step, 1 - h["betas"][1] ** step])
    pc, gc, mc, vc = (t.clone().cuda() for t in (p, g, m, v))
    shadow = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    seg = (C.c_int64 * 3)(*ends)
    ops.adamw(pc, gc, mc, vc, shadow, hyper.cuda(), seg, zero_grad=True)
    lo = 0
    for i, hi in enumerate(ends):
        adamw_step(p[lo:hi], g[lo:hi], m[lo:hi], v[lo:hi], step, **hp[i])
        lo = hi
    assert rel(pc.cpu(), p) < 1e-6 and rel(mc.cpu(), m) < 1e-6 and rel(vc.cpu(), v) < 1e-6
    assert gc.abs().max() == 0
    assert torch.equal(shadow.cpu(), p.to(torch.bfloat16)) or rel(shadow.cpu(), p) < 1e-2


# ---------------------------------------------------------------------------------------------------------------------
# bi-mask gates + architecture losses (layers.py:494-509, 847-858, 179-191; base_model.py:31-86)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("D,H,depth,dead", [(192, 3, 2, False), (384, 6, 3, True), (768, 12, 1, False)])
def test_bimask_fwd_bwd_vs_oracle(cuda_dev, D, H, depth, dead):
    from ofb_b200.engine import BimaskTable
    from fixtures import make_params
    from ofb_oracle import (ModelCfg, _sparsity_term, default_switches, embed_widths, gate_1d, gate_attn,
                            head_channel_widths, head_counts, hidden_widths)
    cfg = ModelCfg(embed_dim=D, num_heads=H, depth=depth)
    P = make_params(cfg, seed=3)
    sw = default_switches(cfg)
    if dead:
        g = torch.Generator().manual_seed(5)
        for k, s in sw.items():
            kill = torch.rand(s.shape, generator=g) < 0.3
            kill.view(-1)[0] = False
            kill.view(-1)[-1] = False
            sw[k] = ~kill
    names = [k for k in P if k.endswith(".alpha") or k.endswith(".score")]
    w_p = 0.7
    tab = BimaskTable(cfg.embed_dim, cfg.num_heads, cfg.depth, cfg.hidden, sw, w_attn=0.5, w_mlp=0.5, w_embed=0.5,
                      w_flops=5.0, target_flops=1.0, num_classes=1000, num_patches=196)
    offs, flat, o = {}, [], 0
    for k in names:
        offs[k] = o
        flat.append(P[k].reshape(-1))
        o += P[k].numel()
    params = torch.cat(flat).cuda()
    tab.bind(offs, params.device)
    wp_dev = torch.tensor([w_p], device="cuda")
    tab.forward(params, wp_dev)
    torch.cuda.synchronize()
    # ---- pruned-unit index sets after thresholding: must match the reference's argsort rule EXACTLY ----
    from ofb_oracle import keep_index_sets
    got_sets = tab.pruned_index_sets()
    for mod in tab.modules:
        pre = mod["prefix"]
        sc = P[pre + ".score"]
        if mod["kind"] == 2:
            ref_sets = keep_index_sets(sc, head_channel_widths(cfg.head_dim), head_counts(cfg.num_heads))
        else:
            ref_sets = keep_index_sets(sc, embed_widths(D) if mod["kind"] == 0 else hidden_widths(cfg.hidden))
        assert got_sets[pre] == ref_sets, pre
    # ---- oracle ----
    leaves = {k: P[k].clone().requires_grad_(True) for k in names}
    gates, wsums, terms = {}, {}, {"attn": 0., "mlp": 0., "embed": 0.}
    for mod in tab.modules:
        pre = mod["prefix"]
        a, s = leaves[pre + ".alpha"], leaves[pre + ".score"]
        if mod["kind"] == 2:
            gt, _, ws = gate_attn(a, sw[pre], s, head_counts(H), head_channel_widths(64), w_p)
            coef, key = 4e-4, "attn"
        elif mod["kind"] == 1:
            gt, _, ws = gate_1d(a, sw[pre], s, hidden_widths(cfg.hidden), w_p)
            coef, key = 1e-4, "mlp"
        else:
            gt, _, ws = gate_1d(a, sw[pre], s, embed_widths(D), w_p)
            coef, key = 1e-4, "embed"
        gates[pre], wsums[pre] = gt, ws
        if int(sw[pre].sum()) > 1:
            terms[key] = terms[key] + _sparsity_term(a, sw[pre], s, coef)
        got = tab.gate[mod["gate_off"]:mod["gate_off"] + gt.numel()].cpu()
        assert rel(got, gt.detach().reshape(-1)) < FP32_TOL, pre
    # FLOPs loss through the oracle's full forward is covered by the step test; here: same polynomial
    from ofb_b200.engine import flops_polynomial
    f_ori, f_s = flops_polynomial(cfg.embed_dim, H, 64, cfg.hidden, 196, 1000, depth, wsums["patch_embed"],
                                  [wsums[f"blocks.{l}.attn"] for l in range(depth)],
                                  [wsums[f"blocks.{l}.mlp"] for l in range(depth)])
    l_flops = ((f_s / 1e9 - 1.0) / (f_ori / 1e9)) ** 2
    arch = 0.5 * terms["attn"] + 0.5 * terms["mlp"] + 0.5 * terms["embed"] + 5.0 * l_flops
    got_arch = tab.arch.cpu()
    assert rel(got_arch[0], arch.detach()) < FP32_TOL
    assert rel(got_arch[4], l_flops.detach()) < 1e-3
    # ---- backward: random upstream d gate ----
    torch.manual_seed(1)
    dgate = torch.randn(tab.total_gate, device="cuda") * 0.1
    up = sum((gates[m["prefix"]].reshape(-1) * dgate[m["gate_off"]:m["gate_off"] + gates[m["prefix"]].numel()].cpu()).sum()
             for m in tab.modules)
    (arch + up).backward()
    grads = torch.zeros_like(params)
    tab.backward(params, wp_dev, dgate, 1.0, grads)
    torch.cuda.synchronize()
    for k in names:
        got = grads[offs[k]:offs[k] + P[k].numel()].cpu()
        want = leaves[k].grad.reshape(-1)
        assert rel(got, want) < 2e-4, k


def test_hyper_uploader_ring(cuda_dev):
    """ofb_copy_f32 reads a pinned HOST slot directly (no copy engine); the ring hands out a fresh slot per step so that values of
    an enqueued upload are never overwritten by a host running ahead."""
    from ofb_b200 import ops
    up = ops.HyperUploader(300, "cuda", slots=3)
    dst = torch.zeros(300, device="cuda")
    seen = []
    for step in range(7):
        h = up.begin()
        h.copy_(torch.arange(300, dtype=torch.float32) + 1000 * step)
        up.upload(dst)
        seen.append(dst.clone())                       # stream-ordered snapshot, no host sync between steps
    h = up.begin(keep=True)
    assert float(h[5]) == 6005.0
    torch.cuda.synchronize()
    for step, t in enumerate(seen):
        assert torch.equal(t.cpu(), torch.arange(300, dtype=torch.float32) + 1000 * step)
