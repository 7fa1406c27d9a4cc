"""GPU parity tests of every kernel behind the C ABI against plain fp32 torch on the same
bf16-rounded inputs (per-kernel tolerance: <= ~1 bf16 ulp of the output scale)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16 = torch.bfloat16


def _ops():
    from t2v_turbo_b200 import ops
    return ops


def rnd(*shape, scale=1.0, seed=0, device="cuda"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(device)


def assert_close(got, ref, rtol=8e-3, atol_scale=4e-3, what=""):
    """|got - ref| <= atol_scale * max|ref| + rtol * |ref| elementwise: 2 bf16 ulps of the element plus 1 bf16 ulp of the
    output scale (round 2: half of round 1's bound; the observed worst err / tol ratio is printed with -s and stays
    below 0.5 for every kernel, i.e. the bound is ~2x the observed error)."""
    got = got.float()
    ref = ref.float()
    scale = ref.abs().max().item() + 1e-6
    err = (got - ref).abs()
    tol = atol_scale * scale + rtol * ref.abs()
    bad = err > tol
    print(f"[tol] {what}: worst err/tol {(err / tol).max().item():.3f}, max err {err.max().item():.3e}, scale {scale:.3e}")
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.numel()} mismatches, max err {err.max().item():.4g} "
                           f"(scale {scale:.4g}), first bad idx {bad.nonzero()[0].tolist()}")
    assert torch.isfinite(got).all(), f"{what}: non-finite output"


# ----------------------------------------------------------------------------- GEMM geometries
@pytest.mark.parametrize("m,k,n,bn", [
    (256, 64, 64, 64), (1000, 320, 320, 0), (640, 1280, 1280, 128), (130, 128, 160, 160),
    (2560, 320, 2560, 256), (77, 1024, 640, 0), (4096, 512, 512, 0),
])
def test_linear(cuda_device, m, k, n, bn):
    ops = _ops()
    x = rnd(m, k, seed=1).to(BF16)
    w = rnd(n, k, scale=k ** -0.5, seed=2).to(BF16)
    b = rnd(n, seed=3)
    res = rnd(m, n, seed=4).to(BF16)
    out = ops.linear(x, w, b, residual=res, block_n=bn)
    ref = x.float() @ w.float().t() + b + res.float()
    assert_close(out, ref, what=f"linear {m}x{k}x{n}")
    out2 = ops.linear(x, w, None, block_n=bn)
    assert_close(out2, x.float() @ w.float().t(), what="linear nobias")


@pytest.mark.parametrize("m,k,n,split", [(640, 1280, 1280, 4), (640, 5120, 1280, 0), (2560, 2560, 320, 3), (77, 1024, 2560, 0)])
def test_linear_split_k(cuda_device, m, k, n, split):
    ops = _ops()
    x = rnd(m, k, seed=70).to(BF16)
    w = rnd(n, k, scale=k ** -0.5, seed=71).to(BF16)
    b = rnd(n, seed=72)
    res = rnd(m, n, seed=73).to(BF16)
    out = ops.linear(x, w, b, residual=res, split_k=split)
    assert_close(out, x.float() @ w.float().t() + b + res.float(), what=f"linear split_k={split}")


def test_conv3x3_level3_split_k(cuda_device):
    """the 5x8 level: M = 640 points, K = 9*1280 -> automatic split-K"""
    ops = _ops()
    x = rnd(16, 5, 8, 1280, seed=74).to(BF16)
    wt = rnd(1280, 1280, 3, 3, scale=(9 * 1280) ** -0.5, seed=75).to(BF16)
    b = rnd(1, 1280, seed=76)
    res = rnd(16, 5, 8, 1280, seed=77).to(BF16)
    out = ops.conv3x3(x, ops.pack_conv_weight(wt), b, bias_div=16, residual=res)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), b[0], padding=1).permute(0, 2, 3, 1) + res.float()
    assert_close(out, ref, what="conv3x3 5x8 split-K")


@pytest.mark.parametrize("m,k,n,bn,geglu", [
    (9600 + 64, 320, 320, 0, False),      # odd number of 128-row tiles (+ ragged tail): last pair has an empty half
    (16384, 640, 1280, 256, False), (16384, 320, 640, 128, False), (12800, 1280, 320, 160, False),
    (16384, 320, 2560, 256, True), (20480, 64, 512, 128, True),
])
def test_linear_cta_pairs(cuda_device, m, k, n, bn, geglu):
    """shapes large enough for the cta_group::2 path (>= 74 pair tiles)"""
    ops = _ops()
    x = rnd(m, k, seed=80).to(BF16)
    w = rnd(n, k, scale=k ** -0.5, seed=81).to(BF16)
    b = rnd(n, seed=82)
    if geglu:
        wp, bp = ops.pack_geglu(w, b.to(BF16))
        out = ops.linear(x, wp, bp, geglu=True, block_n=bn)
        h = x.float() @ w.float().t() + b.to(BF16).float()
        a, g = h.chunk(2, dim=-1)
        ref = a * F.gelu(g)
    else:
        res = rnd(m, n, seed=83).to(BF16)
        out = ops.linear(x, w, b, residual=res, block_n=bn)
        ref = x.float() @ w.float().t() + b + res.float()
    assert_close(out, ref, what=f"linear pair {m}x{k}x{n}")


def test_conv_cta_pairs(cuda_device):
    ops = _ops()
    x = rnd(16, 40, 64, 64, seed=84).to(BF16)
    wt = rnd(320, 64, 3, 3, scale=(9 * 64) ** -0.5, seed=85).to(BF16)
    b = rnd(2, 320, seed=86)
    res = rnd(16, 40, 64, 320, seed=87).to(BF16)
    out = ops.conv3x3(x, ops.pack_conv_weight(wt), b, bias_div=8, residual=res)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), None, padding=1).permute(0, 2, 3, 1) + res.float()
    ref = ref + b.repeat_interleave(8, 0)[:, None, None, :]
    assert_close(out, ref, what="conv3x3 pair")
    xt = rnd(1, 16, 2560, 64, seed=88).to(BF16)
    wtt = rnd(128, 64, 3, 1, 1, scale=(3 * 64) ** -0.5, seed=89).to(BF16)
    bt = rnd(128, seed=90)
    out = ops.tconv3(xt, ops.pack_conv_weight(wtt), bt)
    ref = F.conv3d(xt.float().permute(0, 3, 1, 2).unsqueeze(-1), wtt.float(), bt, padding=(1, 0, 0)).squeeze(-1).permute(0, 2, 3, 1)
    assert_close(out, ref, what="tconv3 pair")


def test_linear_f32_out_and_tail(cuda_device):
    ops = _ops()
    x = rnd(300, 192, seed=5).to(BF16)
    w = rnd(4, 192, scale=0.1, seed=6).to(BF16)   # N = 4 (UNet out conv width)
    b = rnd(4, seed=7)
    out = ops.linear(x, w, b)
    assert out.shape == (300, 4)
    assert_close(out, x.float() @ w.float().t() + b, what="linear N=4")
    out = ops.linear(x, w, b, out_f32=True)
    assert out.dtype == torch.float32
    assert_close(out, x.float() @ w.float().t() + b, rtol=2e-3, atol_scale=1e-3, what="linear f32")


def test_linear_dual_source(cuda_device):
    ops = _ops()
    xa = rnd(500, 128, seed=8).to(BF16)
    xb = rnd(500, 64, seed=9).to(BF16)
    w = rnd(320, 192, scale=0.08, seed=10).to(BF16)
    out = ops.linear((xa, xb), w, None)
    ref = torch.cat([xa, xb], 1).float() @ w.float().t()
    assert_close(out, ref, what="linear dual")


@pytest.mark.parametrize("m,c", [(512, 320), (1000, 64)])
def test_geglu(cuda_device, m, c):
    ops = _ops()
    inner = 4 * c
    x = rnd(m, c, seed=11).to(BF16)
    w = rnd(2 * inner, c, scale=c ** -0.5, seed=12).to(BF16)
    b = rnd(2 * inner, seed=13, scale=0.5).to(BF16)
    wp, bp = ops.pack_geglu(w, b)
    out = ops.linear(x, wp, bp, geglu=True)
    h = x.float() @ w.float().t() + b.float()
    a, g = h.chunk(2, dim=-1)
    ref = a * F.gelu(g)
    assert out.shape == (m, inner)
    assert_close(out, ref, what="geglu")


# LayerNorm folded into the consuming Linear (attention.py:279-281 -> to_q/to_k/to_v, GEGLU.proj): the GEMM reads the
# un-normalised x, the epilogue applies rstd / mean per row.  Reference: LN then Linear in fp32.
@pytest.mark.parametrize("m,c,n,bias,split_k", [(1000, 320, 960, False, 0), (2560, 1280, 1280, True, 0), (640, 1280, 1280, False, 0),
                                                 (130, 640, 320, True, 4), (77, 320, 320, True, 0)])
def test_linear_folded_layernorm(cuda_device, m, c, n, bias, split_k):
    ops = _ops()
    x = (rnd(m, c, seed=51) * 1.7 + 0.8).to(BF16)          # non-zero mean: the mean correction matters
    w = rnd(n, c, scale=c ** -0.5, seed=52)
    b = rnd(n, seed=53) if bias else None
    g = rnd(c, seed=54) * 0.3 + 1
    be = rnd(c, seed=55) * 0.3
    wp, bp, cs = ops.fold_layernorm(w, b, g, be)
    st = ops.layernorm_stats(x, 1e-5)
    xf = x.float()
    mu, var = xf.mean(1), xf.var(1, unbiased=False)
    torch.testing.assert_close(st[:, 0], (var + 1e-5).rsqrt(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(st[:, 1], -(var + 1e-5).rsqrt() * mu, rtol=1e-4, atol=1e-5)
    out = ops.linear(x, wp, bp, ln=(st, cs), split_k=split_k)
    ref = F.linear(F.layer_norm(xf, (c,), g, be, 1e-5), w, b)
    assert_close(out, ref, what="linear + folded layernorm")


@pytest.mark.parametrize("m,k,c,n,split_k", [(1000, 320, 320, 960, 0), (2560, 1280, 1280, 1280, 0), (640, 1280, 1280, 3840, 0),
                                              (300, 2048, 640, 640, 4)])
def test_layernorm_stats_from_producer_gemm(cuda_device, m, k, c, n, split_k):
    """No LayerNorm kernel at all: GEMM 1 (out-projection + residual) accumulates per-row sum / sum of squares of its
    output in the epilogue, GEMM 2 consumes them (raw) with the LayerNorm folded into its weights."""
    ops = _ops()
    a = rnd(m, k, seed=61).to(BF16)
    w1 = rnd(c, k, scale=k ** -0.5, seed=62).to(BF16)
    b1 = rnd(c, seed=63)
    res = (rnd(m, c, seed=64) + 0.4).to(BF16)
    acc = torch.zeros(m, 2, device=a.device, dtype=torch.float32)
    x = ops.linear(a, w1, b1, residual=res, row_accum=acc, split_k=split_k)
    xf = x.float()
    torch.testing.assert_close(acc[:, 0], xf.sum(1), rtol=2e-3, atol=0.15)       # sums of the pre-rounding fp32 values
    torch.testing.assert_close(acc[:, 1], (xf * xf).sum(1), rtol=4e-3, atol=0.15)
    w2 = rnd(n, c, scale=c ** -0.5, seed=65)
    g = rnd(c, seed=66) * 0.3 + 1
    be = rnd(c, seed=67) * 0.3
    wp, bp, cs = ops.fold_layernorm(w2, None, g, be)
    out = ops.linear(x, wp, bp, ln=(acc, cs, (c, 1e-5)))
    ref = F.linear(F.layer_norm(xf, (c,), g, be, 1e-5), w2)
    assert_close(out, ref, what="producer-accumulated layernorm")


@pytest.mark.parametrize("m,c", [(1000, 320), (256, 640)])
def test_geglu_folded_layernorm(cuda_device, m, c):
    ops = _ops()
    inner = 4 * c
    x = (rnd(m, c, seed=56) * 1.5 - 0.6).to(BF16)
    w = rnd(2 * inner, c, scale=c ** -0.5, seed=57)
    b = rnd(2 * inner, seed=58, scale=0.5)
    g = rnd(c, seed=59) * 0.3 + 1
    be = rnd(c, seed=60) * 0.3
    wf, bf, cs = ops.fold_layernorm(w, b, g, be)
    wp, bp = ops.pack_geglu(wf, bf)
    _, csp = ops.pack_geglu(wf, cs)
    out = ops.linear(x, wp, bp, geglu=True, ln=(ops.layernorm_stats(x, 1e-5), csp))
    h = F.linear(F.layer_norm(x.float(), (c,), g, be, 1e-5), w, b)
    a, gate = h.chunk(2, dim=-1)
    assert_close(out, a * F.gelu(gate), what="geglu + folded layernorm")


@pytest.mark.parametrize("n,h,w,cin,cout", [
    (2, 8, 16, 64, 64), (3, 10, 16, 128, 192), (2, 5, 8, 128, 64), (2, 40, 64, 64, 320), (1, 20, 32, 320, 128),
])
def test_conv3x3(cuda_device, n, h, w, cin, cout):
    ops = _ops()
    x = rnd(n, h, w, cin, seed=14).to(BF16)
    wt = rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=15).to(BF16)
    b = rnd(cout, seed=16)
    res = rnd(n, h, w, cout, seed=17).to(BF16)
    wp = ops.pack_conv_weight(wt)
    out = ops.conv3x3(x, wp, b.view(1, -1), bias_div=n, residual=res)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), b, padding=1).permute(0, 2, 3, 1) + res.float()
    assert_close(out, ref, what=f"conv3x3 {n}x{h}x{w} {cin}->{cout}")


def test_conv3x3_rowbias_and_concat(cuda_device):
    ops = _ops()
    bsz, t, h, w = 2, 3, 10, 16
    xa = rnd(bsz * t, h, w, 128, seed=18).to(BF16)
    xb = rnd(bsz * t, h, w, 64, seed=19).to(BF16)
    wt = rnd(128, 192, 3, 3, scale=(9 * 192) ** -0.5, seed=20).to(BF16)
    bias_rows = rnd(bsz, 128, seed=21)   # per-video bias row (timestep embedding add)
    wp = ops.pack_conv_weight(wt)
    out = ops.conv3x3((xa, xb), wp, bias_rows, bias_div=t)
    xin = torch.cat([xa, xb], -1).float().permute(0, 3, 1, 2)
    ref = F.conv2d(xin, wt.float(), None, padding=1).permute(0, 2, 3, 1)
    ref = ref + bias_rows.repeat_interleave(t, 0)[:, None, None, :]
    assert_close(out, ref, what="conv3x3 concat+rowbias")


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 8, 16, 64, 64), (3, 20, 32, 128, 128), (2, 10, 16, 64, 192)])
def test_conv3x3_s2(cuda_device, n, h, w, cin, cout):
    ops = _ops()
    x = rnd(n, h, w, cin, seed=22).to(BF16)
    wt = rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=23).to(BF16)
    b = rnd(cout, seed=24)
    out = ops.conv3x3_s2(x, ops.pack_conv_weight(wt), b)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), b, stride=2, padding=1).permute(0, 2, 3, 1)
    assert_close(out, ref, what="conv3x3 s2")


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 8, 16, 64, 64), (3, 20, 32, 128, 64), (2, 5, 8, 128, 256), (1, 40, 64, 64, 128)])
def test_upconv3x3(cuda_device, n, h, w, cin, cout):
    """nearest-2x upsample + 3x3 conv as four pre-summed 2x2 convs on the low-resolution input
    (reference: Upsample.forward, openaimodel3d.py:96-108 / ae_modules.py:58-63)."""
    ops = _ops()
    x = rnd(n, h, w, cin, seed=29).to(BF16)
    wt = rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=30).to(BF16)
    b = rnd(cout, seed=31)
    out = ops.upconv3x3(x, ops.pack_upconv_weight(wt), b)
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    ref = F.conv2d(up, wt.float(), b, padding=1).permute(0, 2, 3, 1)
    assert out.shape == ref.shape
    assert_close(out, ref, what="upconv3x3")


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 8, 16, 64, 64), (3, 20, 32, 128, 128), (1, 40, 64, 128, 128)])
def test_conv3x3_s2_pad_bottom_right(cuda_device, n, h, w, cin, cout):
    """KL-VAE Downsample (ae_modules.py:87-105): F.pad(x, (0, 1, 0, 1)) then a stride-2 3x3 conv without padding."""
    ops = _ops()
    x = rnd(n, h, w, cin, seed=97).to(BF16)
    wt = rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=98).to(BF16)
    b = rnd(cout, seed=99)
    out = ops.conv3x3_s2(x, ops.pack_conv_weight(wt), b, pad="br")
    xp = F.pad(x.float().permute(0, 3, 1, 2), (0, 1, 0, 1))
    ref = F.conv2d(xp, wt.float(), b, stride=2).permute(0, 2, 3, 1)
    assert out.shape == ref.shape
    assert_close(out, ref, what="conv3x3 s2 (pad bottom/right)")


@pytest.mark.parametrize("dtype", [torch.float32, BF16])
def test_gaussian_sample(cuda_device, dtype):
    """distributions.py:24-42 + ddpm3d.py:558-567: scale * (mean + exp(0.5 clamp(logvar)) * noise), and the mode."""
    ops = _ops()
    b, t, h, w, zc = 2, 3, 5, 8, 4
    mom = (rnd(b * t, h, w, 2 * zc, seed=100) * 20).float().contiguous()     # logvar beyond the clamp range too
    noise = rnd(b * t, zc, h, w, seed=101).float()
    z = ops.gaussian_sample(mom, noise, b=b, t=t, zc=zc, scale=0.18215, dtype=dtype)
    zm = ops.gaussian_sample(mom, None, b=b, t=t, zc=zc, scale=0.18215, dtype=dtype)
    m = mom.permute(0, 3, 1, 2)
    mean, logvar = m[:, :zc], torch.clamp(m[:, zc:], -30.0, 20.0)
    ref = 0.18215 * (mean + torch.exp(0.5 * logvar) * noise)
    back = lambda v: v.reshape(b, t, zc, h, w).permute(0, 2, 1, 3, 4)
    tol = dict(rtol=1e-5, atol=1e-5) if dtype == torch.float32 else dict(rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(z.float(), back(ref), **tol)
    torch.testing.assert_close(zm.float(), back(0.18215 * mean), **tol)


@pytest.mark.parametrize("b,t,hw,c", [(1, 16, 40, 128), (2, 4, 160, 64), (1, 16, 640, 64)])
def test_tconv3(cuda_device, b, t, hw, c):
    ops = _ops()
    x = rnd(b, t, hw, c, seed=25).to(BF16)
    wt = rnd(c, c, 3, 1, 1, scale=(3 * c) ** -0.5, seed=26).to(BF16)
    bias = rnd(c, seed=27)
    res = rnd(b, t, hw, c, seed=28).to(BF16)
    out = ops.tconv3(x, ops.pack_conv_weight(wt), bias, residual=res)
    xin = x.float().permute(0, 3, 1, 2).unsqueeze(-1)  # b c t hw 1
    ref = F.conv3d(xin, wt.float(), bias, padding=(1, 0, 0)).squeeze(-1).permute(0, 2, 3, 1) + res.float()
    assert_close(out, ref, what="tconv3")


def test_bmm_nt(cuda_device):
    ops = _ops()
    a = rnd(3, 200, 128, seed=29).to(BF16)
    b = rnd(3, 136, 128, seed=30).to(BF16)
    out = ops.bmm_nt(a, b, alpha=0.5)
    ref = 0.5 * torch.einsum("bmk,bnk->bmn", a.float(), b.float())
    assert_close(out, ref, what="bmm_nt")


@pytest.mark.parametrize("n,h,w,cout", [(3, 10, 16, 64), (2, 7, 10, 64), (16, 40, 64, 320), (2, 40, 64, 512), (1, 5, 4, 8)])
def test_conv_small_cin(cuda_device, n, h, w, cout):
    """4-channel latent conv: the 4-pixels-per-thread path (width % 4 == 0) and the one-pixel path (ragged width)."""
    ops = _ops()
    x = rnd(n, h, w, 4, seed=31).to(BF16)
    wt = rnd(cout, 4, 3, 3, scale=1 / 6, seed=32).to(BF16)
    b = rnd(cout, seed=33)
    out = ops.conv3x3_small_cin(x, ops.pack_conv_weight(wt), b, cout)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), b, padding=1).permute(0, 2, 3, 1)
    assert_close(out, ref, what="conv small cin")


# ----------------------------------------------------------------------------- norms
# GroupNorm statistics accumulated by the producing GEMM's epilogue (col_accum) and consumed by t2v_groupnorm (chan_sums)
def _chan_sums_ref(y, n_samples):
    yf = y.float().reshape(n_samples, -1, y.shape[-1])
    return torch.stack([yf.sum(1), (yf * yf).sum(1)], dim=-1)


@pytest.mark.parametrize("n,h,w,cin,cout", [(4, 8, 16, 64, 64), (16, 10, 16, 128, 320), (16, 5, 8, 128, 128), (2, 40, 64, 64, 320),
                                             (3, 20, 32, 64, 640), (2, 36, 56, 64, 128), (5, 16, 24, 128, 256)])
def test_conv3x3_groupnorm_stats(cuda_device, n, h, w, cin, cout):
    ops = _ops()
    x = rnd(n, h, w, cin, seed=70).to(BF16)
    wt = rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=71).to(BF16)
    b = rnd(1, cout, seed=72)
    res = (rnd(n, h, w, cout, seed=73) + 0.3).to(BF16)
    stats = torch.zeros(n, cout, 2, device=x.device)
    y = ops.conv3x3(x, ops.pack_conv_weight(wt), b, bias_div=n, residual=res, stats=stats)
    ref = _chan_sums_ref(y, n)
    torch.testing.assert_close(stats, ref, rtol=2e-3, atol=2e-2)
    g = rnd(cout, seed=74) * 0.2 + 1
    be = rnd(cout, seed=75) * 0.2
    out = ops.groupnorm(y.view(-1, cout), g, be, rows_per_sample=h * w, eps=1e-5, silu=True, chan_sums=stats)
    gn = F.silu(F.group_norm(y.float().permute(0, 3, 1, 2), 32, g, be, 1e-5)).permute(0, 2, 3, 1).reshape(-1, cout)
    assert_close(out, gn, what="groupnorm from producer sums")


@pytest.mark.parametrize("n,h,w,c0,c1,cout,t", [(8, 8, 8, 64, 64, 64, 4), (4, 4, 4, 64, 0, 128, 4), (16, 40, 64, 64, 64, 320, 16)])
def test_conv3x3_groupnorm_stats_no_residual(cuda_device, n, h, w, c0, c1, cout, t):
    """ResBlock in_layers conv: concatenated input, per-batch-element bias rows (emb), no residual."""
    ops = _ops()
    xa = rnd(n, h, w, c0, seed=90).to(BF16)
    xb = rnd(n, h, w, c1, seed=91).to(BF16) if c1 else None
    cin = c0 + c1
    wt = rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=92).to(BF16)
    bias = rnd(n // t, cout, seed=93)
    stats = torch.zeros(n, cout, 2, device=xa.device)
    y = ops.conv3x3((xa, xb) if xb is not None else xa, ops.pack_conv_weight(wt), bias, bias_div=t, stats=stats)
    xin = torch.cat([xa, xb], -1) if xb is not None else xa
    ref = F.conv2d(xin.float().permute(0, 3, 1, 2), wt.float(), None, padding=1).permute(0, 2, 3, 1) + bias.repeat_interleave(t, 0)[:, None, None, :]
    assert_close(y, ref, what="conv3x3 (stats, no residual)")
    torch.testing.assert_close(stats, _chan_sums_ref(y, n), rtol=2e-3, atol=2e-2)


@pytest.mark.parametrize("n,h,w,cin,cout", [(8, 16, 16, 64, 64), (4, 8, 8, 64, 128), (8, 32, 32, 64, 64), (16, 40, 64, 64, 320), (3, 80, 128, 64, 128)])
def test_conv3x3_s2_groupnorm_stats(cuda_device, n, h, w, cin, cout):
    ops = _ops()
    x = rnd(n, h, w, cin, seed=94).to(BF16)
    wt = rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=95).to(BF16)
    stats = torch.zeros(n, cout, 2, device=x.device)
    y = ops.conv3x3_s2(x, ops.pack_conv_weight(wt), rnd(cout, seed=96), stats=stats)
    torch.testing.assert_close(stats, _chan_sums_ref(y, n), rtol=2e-3, atol=2e-2)


@pytest.mark.parametrize("b,t,hw,c", [(1, 16, 40, 128), (2, 4, 160, 64), (1, 16, 128, 128), (2, 4, 640, 64), (1, 3, 2560, 64)])
def test_tconv3_groupnorm_stats_temporal(cuda_device, b, t, hw, c):
    """per-frame sums from the (3,1,1) conv, consumed by a GroupNorm over (t, hw) per batch element (chan_group = t)."""
    ops = _ops()
    x = rnd(b, t, hw, c, seed=76).to(BF16)
    wt = rnd(c, c, 3, 1, 1, scale=(3 * c) ** -0.5, seed=77).to(BF16)
    stats = torch.zeros(b * t, c, 2, device=x.device)
    y = ops.tconv3(x, ops.pack_conv_weight(wt), rnd(c, seed=78), stats=stats)
    torch.testing.assert_close(stats, _chan_sums_ref(y, b * t), rtol=2e-3, atol=2e-2)
    g = rnd(c, seed=79) * 0.2 + 1
    be = rnd(c, seed=80) * 0.2
    out = ops.groupnorm(y.view(-1, c), g, be, rows_per_sample=t * hw, eps=1e-5, silu=True, chan_sums=stats, chan_group=t)
    gn = F.silu(F.group_norm(y.float().view(b, t * hw, c).permute(0, 2, 1), 32, g, be, 1e-5)).permute(0, 2, 1).reshape(-1, c)
    assert_close(out, gn, what="temporal groupnorm from per-frame sums")


@pytest.mark.parametrize("nf,hw,k,n,split_k", [(16, 160, 320, 320, 0), (16, 40, 1280, 1280, 0), (16, 256, 320, 320, 0), (6, 2560, 320, 320, 0),
                                               (4, 640, 128, 640, 0), (16, 40, 2048, 128, 4)])
def test_linear_frames_groupnorm_stats_concat(cuda_device, nf, hw, k, n, split_k):
    """proj_out (+residual) tiled per frame; its sums feed a GroupNorm over the concatenation with a second tensor."""
    ops = _ops()
    x = rnd(nf * hw, k, seed=81).to(BF16)
    w = rnd(n, k, scale=k ** -0.5, seed=82).to(BF16)
    res = rnd(nf * hw, n, seed=83).to(BF16)
    stats = torch.zeros(nf, n, 2, device=x.device)
    y = ops.linear_frames(x, w, rnd(n, seed=84), hw=hw, residual=res, stats=stats, split_k=split_k)
    assert_close(y, x.float() @ w.float().t() + rnd(n, seed=84) + res.float(), what="linear_frames")
    torch.testing.assert_close(stats, _chan_sums_ref(y, nf), rtol=2e-3, atol=2e-2)
    skip = (rnd(nf * hw, 64, seed=85) * 2).to(BF16)
    s2 = _chan_sums_ref(skip, nf).contiguous()
    c = n + 64
    g = rnd(c, seed=86) * 0.2 + 1
    be = rnd(c, seed=87) * 0.2
    out = ops.groupnorm((y, skip), g, be, rows_per_sample=hw, eps=1e-5, silu=True, chan_sums=(stats, s2))
    cat = torch.cat([y, skip], 1).float().view(nf, hw, c).permute(0, 2, 1)
    gn = F.silu(F.group_norm(cat, 32, g, be, 1e-5)).permute(0, 2, 1).reshape(-1, c)
    assert_close(out, gn, what="concat groupnorm from producer sums")


# mode 1 = statistics + apply kernel pair, mode 2 = single-kernel cluster path (must not fall back), 0 = automatic
@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("n,hw,c,rps_mult,silu", [(4, 160, 320, 1, True), (2, 40, 1280, 2, True), (6, 64, 128, 3, False),
                                                  (16, 2560, 320, 1, True), (3, 77, 64, 1, False), (2, 7, 2560, 1, True)])
def test_groupnorm(cuda_device, n, hw, c, rps_mult, silu, mode):
    ops = _ops()
    x = (rnd(n * hw, c, seed=34) * 2 + 0.5).to(BF16)
    g = rnd(c, seed=35) * 0.2 + 1
    b = rnd(c, seed=36) * 0.2
    out = ops.groupnorm(x, g, b, rows_per_sample=hw * rps_mult, eps=1e-5, silu=silu, mode=mode)
    xs = x.float().view(n // rps_mult, hw * rps_mult, c).permute(0, 2, 1)
    ref = F.group_norm(xs, 32, g, b, 1e-5)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 1).reshape(n * hw, c)
    assert_close(out, ref, what="groupnorm")


def test_groupnorm_too_large_for_cluster(cuda_device):
    """A VAE-sized sample (40960 rows x 128 ch = 10 MB) cannot live in a cluster's shared memory: automatic mode
    uses the two-kernel path, and demanding the cluster path is an error rather than a silent fallback."""
    ops = _ops()
    x = rnd(40960, 128, seed=44).to(BF16)
    g = rnd(128, seed=45) * 0.2 + 1
    b = rnd(128, seed=46) * 0.2
    out = ops.groupnorm(x, g, b, rows_per_sample=40960, eps=1e-6, silu=True)
    ref = F.silu(F.group_norm(x.float().view(1, 40960, 128).permute(0, 2, 1), 32, g, b, 1e-6)).permute(0, 2, 1).reshape(40960, 128)
    assert_close(out, ref, what="groupnorm big")
    with pytest.raises(RuntimeError):
        ops.groupnorm(x, g, b, rows_per_sample=40960, eps=1e-6, silu=True, mode=2)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_groupnorm_concat(cuda_device, mode):
    ops = _ops()
    n, hw = 3, 160
    xa = rnd(n * hw, 1280, seed=37).to(BF16)
    xb = (rnd(n * hw, 640, seed=38) * 3).to(BF16)
    c = 1920
    g = rnd(c, seed=39) * 0.2 + 1
    b = rnd(c, seed=40) * 0.2
    out = ops.groupnorm((xa, xb), g, b, rows_per_sample=hw, eps=1e-5, silu=True, mode=mode)
    xs = torch.cat([xa, xb], 1).float().view(n, hw, c).permute(0, 2, 1)
    ref = F.silu(F.group_norm(xs, 32, g, b, 1e-5)).permute(0, 2, 1).reshape(n * hw, c)
    assert_close(out, ref, what="groupnorm concat")


@pytest.mark.parametrize("rows,c", [(1000, 320), (77, 1280), (513, 512), (64, 640)])
def test_layernorm(cuda_device, rows, c):
    ops = _ops()
    x = (rnd(rows, c, seed=41) * 1.5 + 0.3).to(BF16)
    g = rnd(c, seed=42) * 0.2 + 1
    b = rnd(c, seed=43) * 0.2
    out = ops.layernorm(x, g, b, 1e-5)
    ref = F.layer_norm(x.float(), (c,), g, b, 1e-5)
    assert_close(out, ref, what="layernorm")


# ----------------------------------------------------------------------------- attention
def _sdpa_ref(q, k, v, heads, scale):
    bq, lq, inner = q.shape
    bk, lk, _ = k.shape
    rep = bq // bk
    qh = q.float().view(bq, lq, heads, 64).transpose(1, 2)
    kh = k.float().view(bk, lk, heads, 64).transpose(1, 2).repeat_interleave(rep, 0)
    vh = v.float().view(bk, lk, heads, 64).transpose(1, 2).repeat_interleave(rep, 0)
    p = torch.softmax(qh @ kh.transpose(-1, -2) * scale, dim=-1)
    return (p @ vh).transpose(1, 2).reshape(bq, lq, inner)


@pytest.mark.parametrize("b,lq,lk,heads,rep", [
    (2, 256, 256, 2, 1), (1, 128, 128, 1, 1), (2, 160, 160, 3, 1), (4, 40, 40, 2, 1), (4, 640, 77, 2, 4), (1, 2560, 2560, 5, 1),
    (2, 320, 300, 2, 1), (3, 100, 513, 1, 1), (2, 640, 640, 10, 1), (1, 257, 129, 2, 1),
])
def test_attention(cuda_device, b, lq, lk, heads, rep):
    ops = _ops()
    inner = heads * 64
    q = rnd(b, lq, inner, seed=44).to(BF16)
    k = rnd(b // rep, lk, inner, seed=45).to(BF16)
    v = rnd(b // rep, lk, inner, seed=46).to(BF16)
    out = ops.attention(q, k, v, heads=heads, scale=0.125, kv_batch_div=rep)
    ref = _sdpa_ref(q, k, v, heads, 0.125)
    assert_close(out, ref, what=f"attention b{b} lq{lq} lk{lk} h{heads}")


def test_attention_reference_raise_slow_path(cuda_device):
    """Later key tiles whose scores exceed the first tile's row maximum by far more than 2^100: the exponent reference
    fixed by the first tile must be raised and l / O rescaled exactly (the rare slow path of the two-tile kernel)."""
    ops = _ops()
    b, lq, lk, heads = 2, 256, 512, 2
    inner = heads * 64
    q = rnd(b, lq, inner, seed=144).to(BF16)
    k = rnd(b, lk, inner, seed=145)
    k[:, 300:340] *= 40.0      # tile 2 holds scores ~ +-250 in the log2 domain; tiles 0, 1 stay O(1)
    k[:, 450:452] *= 90.0      # and tile 3 raises the reference a second time for some rows
    k = k.to(BF16)
    v = rnd(b, lk, inner, seed=146).to(BF16)
    out = ops.attention(q, k, v, heads=heads, scale=0.125)
    ref = _sdpa_ref(q, k, v, heads, 0.125)
    assert torch.isfinite(out.float()).all()
    assert_close(out, ref, what="attention, exponent reference raised mid-sequence")


def test_attention_strided_qkv(cuda_device):
    """q/k/v are column slices of one fused projection output (row stride 3*inner)."""
    ops = _ops()
    b, l, heads = 2, 256, 2
    inner = heads * 64
    qkv = rnd(b, l, 3 * inner, seed=47).to(BF16)
    q, k, v = qkv[..., :inner], qkv[..., inner:2 * inner], qkv[..., 2 * inner:]
    out = ops.attention(q, k, v, heads=heads, scale=0.125)
    ref = _sdpa_ref(q.contiguous(), k.contiguous(), v.contiguous(), heads, 0.125)
    assert_close(out, ref, what="attention strided")


@pytest.mark.parametrize("b,t,hw,heads", [(1, 16, 40, 2), (2, 16, 33, 1), (1, 8, 20, 3), (1, 24, 10, 2)])
def test_attention_temporal(cuda_device, b, t, hw, heads):
    ops = _ops()
    inner = heads * 64
    q = rnd(b * t * hw, inner, seed=48).to(BF16)
    k = rnd(b * t * hw, inner, seed=49).to(BF16)
    v = rnd(b * t * hw, inner, seed=50).to(BF16)
    out = ops.attention_temporal(q, k, v, b=b, t=t, hw=hw, heads=heads, scale=0.125)

    def seqs(x):  # (b t hw) c -> (b hw) t c
        return x.view(b, t, hw, inner).permute(0, 2, 1, 3).reshape(b * hw, t, inner)
    ref = _sdpa_ref(seqs(q), seqs(k), seqs(v), heads, 0.125)
    ref = ref.view(b, hw, t, inner).permute(0, 2, 1, 3).reshape(b * t * hw, inner)
    assert_close(out, ref, what="attention temporal")


@pytest.mark.parametrize("b,t,hw,heads,dtype", [(1, 16, 40, 2, torch.bfloat16), (2, 16, 33, 1, torch.float32), (1, 24, 10, 2, torch.float32)])
def test_attention_temporal_prob_export(cuda_device, b, t, hw, heads, dtype):
    """record_attn_probs (attention.py:124-126): softmax(q k^T * scale) in the reference's "(b h) i j" layout."""
    ops = _ops()
    inner = heads * 64
    q = rnd(b * t * hw, inner, seed=148).to(BF16)
    k = rnd(b * t * hw, inner, seed=149).to(BF16)
    v = rnd(b * t * hw, inner, seed=150).to(BF16)
    probs = torch.empty(b * hw * heads, t, t, device="cuda", dtype=dtype)
    out = ops.attention_temporal(q, k, v, b=b, t=t, hw=hw, heads=heads, scale=0.125, probs=probs)
    out2 = ops.attention_temporal(q, k, v, b=b, t=t, hw=hw, heads=heads, scale=0.125)
    assert torch.equal(out, out2)

    def seqs(x):  # (b t hw) (h d) -> ((b hw) h) t d
        return x.view(b, t, hw, heads, 64).permute(0, 2, 3, 1, 4).reshape(b * hw * heads, t, 64).float()
    ref = torch.softmax(seqs(q) @ seqs(k).transpose(1, 2) * 0.125, dim=-1)
    assert_close(probs, ref, rtol=8e-3 if dtype == torch.bfloat16 else 2e-3, atol_scale=4e-3 if dtype == torch.bfloat16 else 1e-3,
                 what=f"temporal attention probabilities t={t}")


# ----------------------------------------------------------------------------- small ops
def test_small_linear_and_sinusoid(cuda_device):
    ops = _ops()
    x = rnd(3, 320, seed=51)
    w = rnd(1280, 320, scale=320 ** -0.5, seed=52).to(BF16)
    b = rnd(1280, seed=53)
    out = ops.small_linear(x, w, b, silu_in=True, silu_out=True, round_bf16=False)
    ref = F.silu(F.silu(x) @ w.float().t() + b)
    assert_close(out, ref, rtol=2e-3, atol_scale=1e-4, what="small_linear")
    t = torch.tensor([999.0, 519.0, 16.0], device="cuda")
    half = 160
    freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half).cuda()
    emb = ops.sinusoidal_embedding(t, freqs, round_bf16=False)
    args = t[:, None] * freqs[None]
    ref = torch.cat([torch.cos(args), torch.sin(args)], -1)
    assert_close(emb, ref, rtol=1e-4, atol_scale=1e-4, what="sinusoid")


def test_layout_and_resample(cuda_device):
    ops = _ops()
    x = rnd(2, 4, 3, 6, 8, seed=54).to(BF16)
    fr = ops.bcthw_to_frames(x, 1.0)
    assert torch.equal(fr, x.permute(0, 2, 3, 4, 1).reshape(6, 6, 8, 4))
    back = ops.frames_to_bcthw(fr, 2, 4, torch.float32)
    assert torch.equal(back, x.float())
    y = rnd(2, 5, 8, 64, seed=55).to(BF16)
    up = ops.upsample_nearest2x(y)
    ref = F.interpolate(y.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up.float(), ref)
    a, bb = rnd(10, 7, 64, seed=56).to(BF16), rnd(10, 7, 128, seed=57).to(BF16)
    assert torch.equal(ops.concat_channels(a, bb), torch.cat([a, bb], -1))
    s = rnd(50, 300, seed=58).to(BF16)
    ref = torch.softmax(s.float() * 0.3, -1)
    assert_close(ops.softmax_rows_(s.clone(), 0.3), ref, what="softmax rows")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_lcm_step(cuda_device, dtype):
    ops = _ops()
    x = rnd(1, 4, 16, 40, 64, seed=59).to(dtype)
    e = rnd(1, 4, 16, 40, 64, seed=60).to(dtype)
    nz = rnd(1, 4, 16, 40, 64, seed=61).to(dtype)
    a_t, a_p = torch.tensor(0.0047), torch.tensor(0.35)
    c_skip, c_out = torch.tensor(2.5e-9), torch.tensor(1.0)
    sb, sa = (1 - a_t).sqrt(), a_t.sqrt()
    x0 = (x - sb.cuda() * e) / sa.cuda()
    den = c_out.cuda() * x0 + c_skip.cuda() * x
    prev = a_p.sqrt().cuda() * den + (1 - a_p).sqrt().cuda() * nz
    p2, d2 = ops.lcm_step(x, e, nz, inv_sqrt_alpha_t=float(1.0 / sa), sqrt_beta_t=float(sb), c_skip=float(c_skip),
                          c_out=float(c_out), sqrt_alpha_prev=float(a_p.sqrt()), sqrt_beta_prev=float((1 - a_p).sqrt()))
    tol = dict(rtol=2e-2, atol_scale=1e-2) if dtype == torch.bfloat16 else dict(rtol=1e-4, atol_scale=1e-5)
    assert_close(d2, den, what="lcm den", **tol)
    assert_close(p2, prev, what="lcm prev", **tol)
