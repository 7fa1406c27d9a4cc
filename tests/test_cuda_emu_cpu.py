"""The SIMT kernels of csrc/train_bwd.cu, train_full.cu, elementwise.cu and train_ops.cu EXECUTED on the CPU: the sources are compiled unchanged by
g++ against tests/cuda_emu/cuda_emu.h (CUDA threads = OS threads, __syncthreads / warp shuffles = barriers, __shared__ = statics,
atomicAdd = std::atomic_ref; the host launch code runs too, with a 2-"SM" device so that grids stay small) and called through
the same C ABI, on CPU tensors.

Two steps: (1) the emulator is validated on kernels that are parity-tested on B200 (t2v_groupnorm_bwd, t2v_layernorm_bwd,
t2v_colsum_samples, t2v_geglu, t2v_resample2x) — if it ran them wrong these tests would fail; (2) the kernels of the v2 full
fine-tune step that have NOT run on a GPU (t2v_groupnorm_affine_grad, t2v_layernorm_affine_grad, t2v_ema_update) are run
the same way against torch.  This executes their indexing, reductions, tails and launch arithmetic; it cannot prove anything about
the hardware (memory ordering, PDL) — the `-m gpu` file tests/test_zz_full_train_gpu.py is there for that."""
import ctypes as C
import os
import shutil
import subprocess

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BF16 = torch.bfloat16
NAMES = ("t2v_groupnorm_bwd", "t2v_layernorm_bwd", "t2v_colsum_samples", "t2v_geglu", "t2v_resample2x", "t2v_ew2d",
         "t2v_groupnorm_affine_grad", "t2v_layernorm_affine_grad", "t2v_ema_update", "t2v_softmax_bwd_rows", "t2v_softmax_rows", "t2v_attn_short_probs_bwd",
         # elementwise.cu / train_ops.cu (all parity-tested on B200; run here as a CPU regression net over the real kernel source)
         "t2v_lcm_step", "t2v_scale_add_rows", "t2v_dropout_scale", "t2v_scale_mask", "t2v_adamw_step", "t2v_sum_squares",
         "t2v_mse_loss_grad", "t2v_huber_loss_grad", "t2v_video_to_uint8", "t2v_conv3x3_small_cin", "t2v_pack_conv_weight",
         "t2v_sinusoidal_embedding", "t2v_bcthw_to_frames_pad", "t2v_frames_to_bcthw", "t2v_concat_channels", "t2v_upsample_nearest2x")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    from t2v_turbo_b200 import _lib
    out = tmp_path_factory.mktemp("cuda_emu") / "libt2v_emu.so"
    csrc = os.path.join(ROOT, "t2v_turbo_b200", "csrc")
    cmd = ["g++", "-std=c++20", "-O1", "-x", "c++", "-DT2V_HOST_EMU", "-I", os.path.join(ROOT, "tests", "cuda_emu"), "-shared", "-fPIC",
           "-pthread"] + [os.path.join(csrc, f) for f in ("train_bwd.cu", "train_full.cu", "elementwise.cu", "train_ops.cu")] + ["-o", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(str(out))
    for n in NAMES:
        fn = getattr(lib, n)
        fn.restype, fn.argtypes = _lib.SYMBOLS[n]
    return lib


def rnd(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def close(got, ref, rtol, atol_scale, what):
    got, ref = got.float(), ref.float()
    tol = atol_scale * (ref.abs().max().item() + 1e-6) + rtol * ref.abs()
    err = (got - ref).abs()
    assert (err <= tol).all(), f"{what}: max err {err.max().item():.4g}, worst err/tol {(err / tol).max().item():.3f}"


def _gn_bwd(emu, x, dy, gamma, beta, hw, eps, silu, add=None):
    from t2v_turbo_b200 import _lib
    rows, c = x.shape
    n = rows // hw
    dx = torch.empty(rows, c, dtype=BF16)
    ws = torch.zeros(n * 32 * 4)
    d = _lib.GroupNormBwdDesc()
    d.x, d.x_row_stride, d.dy, d.dy_row_stride = x.data_ptr(), x.stride(0), dy.data_ptr(), dy.stride(0)
    if add is not None:
        d.dx_add, d.dx_add_row_stride = add.data_ptr(), add.stride(0)
    d.dx, d.dx_row_stride = dx.data_ptr(), dx.stride(0)
    d.gamma, d.beta = gamma.data_ptr(), beta.data_ptr()
    d.rows, d.rows_per_sample, d.channels, d.groups, d.eps, d.silu = rows, hw, c, 32, eps, int(silu)
    d.workspace = ws.data_ptr()
    assert emu.t2v_groupnorm_bwd(C.byref(d), None) == 0
    return dx, ws


def _gn_ref(x, dy, gamma, beta, hw, eps, silu):
    rows, c = x.shape
    n = rows // hw
    xr = x.float().view(n, hw, c).permute(0, 2, 1).contiguous().requires_grad_(True)
    g, b = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = F.group_norm(xr, 32, g, b, eps=eps)
    if silu:
        y = F.silu(y)
    y.backward(dy.float().view(n, hw, c).permute(0, 2, 1))
    return xr.grad.permute(0, 2, 1).reshape(rows, c), g.grad, b.grad


GN_CASES = [(2, 96, 64, True), (3, 50, 320, True), (1, 130, 128, False), (2, 40, 2560, True)]


# ----------------------------------------------------------------------------- (1) the emulator, on GPU-verified kernels
@pytest.mark.parametrize("n,hw,c,silu", GN_CASES)
def test_emulated_groupnorm_bwd_matches_autograd(emu, n, hw, c, silu):
    x = (rnd(n * hw, c, seed=1, scale=1.5) + 0.7).to(BF16)
    dy = rnd(n * hw, c, seed=2).to(BF16)
    gamma, beta = 1.0 + 0.3 * rnd(c, seed=3), 0.2 * rnd(c, seed=4)
    add = rnd(n * hw, c, seed=5).to(BF16)
    dx, _ = _gn_bwd(emu, x, dy, gamma, beta, hw, 1e-5, silu, add)
    ref, _, _ = _gn_ref(x, dy, gamma, beta, hw, 1e-5, silu)
    close(dx, ref + add.float(), 8e-3, 4e-3, f"emulated gn_bwd {n}x{hw}x{c}")        # the B200 test's bf16 bound


@pytest.mark.parametrize("rows,c", [(37, 64), (20, 320), (9, 1280)])
def test_emulated_layernorm_bwd_matches_autograd(emu, rows, c):
    x = (rnd(rows, c, seed=6, scale=2.0) - 0.3).to(BF16)
    dy = rnd(rows, c, seed=7).to(BF16)
    gamma = 1.0 + 0.2 * rnd(c, seed=8)
    dx = torch.empty(rows, c, dtype=BF16)
    assert emu.t2v_layernorm_bwd(x.data_ptr(), c, dy.data_ptr(), c, None, 0, dx.data_ptr(), c, gamma.data_ptr(), rows, c, 1e-5, None) == 0
    xr = x.float().requires_grad_(True)
    F.layer_norm(xr, (c,), gamma, torch.zeros(c), 1e-5).backward(dy.float())
    close(dx, xr.grad, 8e-3, 4e-3, f"emulated ln_bwd {rows}x{c}")


def test_emulated_colsum_geglu_resample(emu):
    x = rnd(6 * 50, 72, seed=9).to(BF16)
    out = torch.full((6, 72), 0.25)
    assert emu.t2v_colsum_samples(x.data_ptr(), 72, out.data_ptr(), 300, 50, 72, None) == 0
    close(out, 0.25 + x.float().view(6, 50, 72).sum(1), 1e-5, 1e-6, "emulated colsum")
    pre = rnd(33, 256, seed=10).to(BF16)
    o = torch.empty(33, 128, dtype=BF16)
    assert emu.t2v_geglu(pre.data_ptr(), 256, None, 0, o.data_ptr(), 128, 33, 128, None) == 0
    close(o, pre.float()[:, :128] * F.gelu(pre.float()[:, 128:]), 8e-3, 4e-3, "emulated geglu")
    img = rnd(2, 6, 8, 16, seed=11).to(BF16)
    pooled = torch.empty(2, 3, 4, 16, dtype=BF16)
    assert emu.t2v_resample2x(2, img.data_ptr(), pooled.data_ptr(), 2, 3, 4, 16, None) == 0
    close(pooled, img.float().view(2, 3, 2, 4, 2, 16).sum((2, 4)), 8e-3, 4e-3, "emulated 2x2 pooling")


# ----------------------------------------------------------------------------- (2) the GPU-unverified kernels of the v2 step
@pytest.mark.parametrize("n,hw,c,silu", GN_CASES)
def test_groupnorm_affine_grad_kernel_under_emulation(emu, n, hw, c, silu):
    x = (rnd(n * hw, c, seed=1, scale=1.5) + 0.7).to(BF16)
    dy = rnd(n * hw, c, seed=2).to(BF16)
    gamma, beta = 1.0 + 0.3 * rnd(c, seed=3), 0.2 * rnd(c, seed=4)
    _, ws = _gn_bwd(emu, x, dy, gamma, beta, hw, 1e-5, silu)          # leaves (sum x, sum x^2) in slots 0 / 1
    dg, db = torch.full((c,), 0.5), torch.full((c,), -0.25)           # accumulated INTO
    assert emu.t2v_groupnorm_affine_grad(x.data_ptr(), c, dy.data_ptr(), c, gamma.data_ptr(), beta.data_ptr(), ws.data_ptr(), dg.data_ptr(),
                                         db.data_ptr(), n * hw, hw, c, 32, 1e-5, int(silu), None) == 0
    _, rg, rb = _gn_ref(x, dy, gamma, beta, hw, 1e-5, silu)
    close(dg, 0.5 + rg, 2e-3, 5e-4, f"gn dgamma {n}x{hw}x{c}")
    close(db, -0.25 + rb, 2e-3, 5e-4, f"gn dbeta {n}x{hw}x{c}")


def test_groupnorm_affine_grad_strided_rows_and_errors(emu):
    n, hw, c = 2, 48, 64
    big, dbig = rnd(n * hw, c + 64, seed=12).to(BF16), rnd(n * hw, c + 128, seed=13).to(BF16)
    x, dy = big[:, 64:], dbig[:, :c]                                    # channel slices of wider tensors: row strides != C
    gamma, beta = 1.0 + 0.1 * rnd(c, seed=14), 0.1 * rnd(c, seed=15)
    _, ws = _gn_bwd(emu, x, dy, gamma, beta, hw, 1e-6, True)
    dg, db = torch.zeros(c), torch.zeros(c)
    assert emu.t2v_groupnorm_affine_grad(x.data_ptr(), x.stride(0), dy.data_ptr(), dy.stride(0), gamma.data_ptr(), beta.data_ptr(),
                                         ws.data_ptr(), dg.data_ptr(), db.data_ptr(), n * hw, hw, c, 32, 1e-6, 1, None) == 0
    _, rg, rb = _gn_ref(x.contiguous(), dy.contiguous(), gamma, beta, hw, 1e-6, True)
    close(dg, rg, 2e-3, 5e-4, "gn dgamma strided")
    close(db, rb, 2e-3, 5e-4, "gn dbeta strided")
    a = (x.data_ptr(), x.stride(0), dy.data_ptr(), dy.stride(0), gamma.data_ptr(), beta.data_ptr(), ws.data_ptr(), dg.data_ptr(), db.data_ptr())
    assert emu.t2v_groupnorm_affine_grad(*a, n * hw, hw + 1, c, 32, 1e-6, 1, None) < 0          # rows % rows_per_sample
    assert emu.t2v_groupnorm_affine_grad(*a, n * hw, hw, 2568, 32, 1e-6, 1, None) < 0           # more channels than the shared table
    assert emu.t2v_groupnorm_affine_grad(x.data_ptr(), 7, *a[2:], n * hw, hw, c, 32, 1e-6, 1, None) < 0   # misaligned row stride


@pytest.mark.parametrize("rows,c", [(37, 64), (20, 320), (300, 128), (9, 1280), (1, 1024)])
def test_layernorm_affine_grad_kernel_under_emulation(emu, rows, c):
    x = (rnd(rows, c, seed=6, scale=2.0) - 0.3).to(BF16)
    dy = rnd(rows, c, seed=7).to(BF16)
    dg, db = torch.ones(c), torch.zeros(c)
    assert emu.t2v_layernorm_affine_grad(x.data_ptr(), c, dy.data_ptr(), c, dg.data_ptr(), db.data_ptr(), rows, c, 1e-5, None) == 0
    xh = F.layer_norm(x.float(), (c,), None, None, 1e-5)
    close(dg, 1.0 + (dy.float() * xh).sum(0), 2e-3, 5e-4, f"ln dgamma {rows}x{c}")
    close(db, dy.float().sum(0), 2e-3, 5e-4, f"ln dbeta {rows}x{c}")
    assert emu.t2v_layernorm_affine_grad(x.data_ptr(), c, dy.data_ptr(), c, dg.data_ptr(), db.data_ptr(), rows, 96, 1e-5, None) < 0


@pytest.mark.parametrize("rows,cols", [(12, 256), (5, 2560), (3, 100), (1, 7)])
def test_softmax_bwd_rows_kernel_under_emulation(emu, rows, cols):
    """The VAE AttnBlock's softmax adjoint (vae.decode WITH grad): forward kernel (B200-verified) and the new backward kernel both run
    under emulation, the backward against autograd of torch.softmax on the forward kernel's own bf16 probabilities."""
    s = (rnd(rows, cols, seed=30) * 3.0).to(BF16)
    scale = 512 ** -0.5
    p = s.clone()
    assert emu.t2v_softmax_rows(p.data_ptr(), rows, cols, cols, scale, None) == 0
    close(p, torch.softmax(s.float() * scale, -1), 8e-3, 4e-3, "emulated softmax_rows")
    dp = rnd(rows, cols, seed=31).to(BF16)
    ds = dp.clone()
    assert emu.t2v_softmax_bwd_rows(ds.data_ptr(), cols, p.data_ptr(), cols, rows, cols, scale, None) == 0
    pf, df = p.float(), dp.float()
    close(ds, scale * pf * (df - (df * pf).sum(-1, keepdim=True)), 8e-3, 4e-3, f"softmax_bwd_rows {rows}x{cols}")
    sr = s.float().requires_grad_(True)                      # and against autograd of the true softmax (bf16 P is the only difference)
    torch.softmax(sr * scale, -1).backward(df)
    assert ((ds.float() - sr.grad).norm() / sr.grad.norm()).item() < 2e-2
    assert emu.t2v_softmax_bwd_rows(ds.data_ptr(), cols - 1, p.data_ptr(), cols, rows, cols, scale, None) < 0


@pytest.mark.parametrize("b,t,hw,heads,strided", [(2, 4, 6, 2, False), (1, 16, 5, 5, True), (1, 3, 9, 1, False)])
def test_attn_short_probs_bwd_kernel_under_emulation(emu, b, t, hw, heads, strided):
    """Adjoint of the temporal attention-probability export w.r.t. q / k (the motion-prior score's path) against autograd of
    softmax(scale q k^T) in the export's "(b hw heads) i j" layout; q / k also as column slices of a fused projection."""
    inner = heads * 64
    rows = b * t * hw
    if strided:
        big = rnd(rows, 3 * inner, seed=32).to(BF16)
        q, k = big[:, :inner], big[:, inner:2 * inner]
    else:
        q, k = rnd(rows, inner, seed=33).to(BF16), rnd(rows, inner, seed=34).to(BF16)
    dp = rnd(b * hw * heads, t, t, seed=35)
    dq, dk = torch.empty(rows, inner, dtype=BF16), torch.empty(rows, inner, dtype=BF16)
    scale = 0.125
    assert emu.t2v_attn_short_probs_bwd(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), dp.data_ptr(), dq.data_ptr(), dk.data_ptr(),
                                        b, hw, heads, t, scale, None) == 0
    qr, kr = q.float().clone().requires_grad_(True), k.float().clone().requires_grad_(True)

    def seqs(x):      # [(b t hw), H*64] -> [(b hw H), t, 64]
        return x.view(b, t, hw, heads, 64).permute(0, 2, 3, 1, 4).reshape(b * hw * heads, t, 64)
    probs = torch.softmax(seqs(qr) @ seqs(kr).transpose(1, 2) * scale, -1)
    (probs * dp).sum().backward()
    close(dq, qr.grad, 8e-3, 4e-3, f"probs_bwd dq b={b} t={t} hw={hw} H={heads}")
    close(dk, kr.grad, 8e-3, 4e-3, f"probs_bwd dk b={b} t={t} hw={hw} H={heads}")
    assert emu.t2v_attn_short_probs_bwd(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), dp.data_ptr(), dq.data_ptr(), dk.data_ptr(),
                                        b, hw, heads, 17, scale, None) < 0


@pytest.mark.parametrize("n,offset", [(4096, 0), (100003, 0), (4099, 1), (3, 0), (1, 0)])
def test_ema_update_kernel_under_emulation(emu, n, offset):
    tgt, src = rnd(n + offset, seed=16)[offset:], rnd(n + offset, seed=17)[offset:]     # offset 1: unaligned for float4 -> scalar path
    ref = tgt.clone().mul_(0.95).add_(src, alpha=0.05)
    assert emu.t2v_ema_update(tgt.data_ptr(), src.data_ptr(), n, 0.95, None) == 0
    assert torch.allclose(tgt, ref, rtol=1e-6, atol=1e-7)
    assert emu.t2v_ema_update(tgt.data_ptr(), src.data_ptr(), n, 1.5, None) < 0


# ----------------------------------------------------------------------------- (3) elementwise.cu / train_ops.cu: a CPU regression net
def test_emulated_dropout_mask_bit_exact_vs_philox_oracle(emu):
    """The in-kernel Philox4x32-10 keep-mask == the numpy oracle (pinned on Random123's known answers), bit for bit — the same
    assertion the B200 suite makes (tests/test_lora_train_gpu.py), here on the kernel source run under emulation."""
    import numpy as np
    from oracle.philox_oracle import keep_mask
    for (rows, c), p, seed, call in (((40, 64), 0.1, 1234, 1), ((24, 64), 0.5, 2 ** 40 + 17, 77), ((3, 8), 0.25, 99, 2 ** 31 + 5)):
        x = rnd(rows, c, seed=18).to(BF16)
        out, keep = torch.empty_like(x), torch.empty(rows, c, dtype=torch.uint8)
        sd = torch.tensor([seed], dtype=torch.int64)
        scale = 1.7 / (1.0 - p)
        assert emu.t2v_dropout_scale(x.data_ptr(), None, out.data_ptr(), keep.data_ptr(), x.numel(), 1.0 - p, scale, sd.data_ptr(), call, None) == 0
        want = keep_mask(x.numel(), 1.0 - p, seed, call).reshape(rows, c)
        assert np.array_equal(keep.numpy(), want)
        assert torch.equal(out, (x.float() * scale * keep.float()).to(BF16))
        dx = torch.empty_like(x)
        assert emu.t2v_scale_mask(x.data_ptr(), keep.data_ptr(), dx.data_ptr(), x.numel(), scale, None) == 0
        assert torch.equal(dx, out)


def test_emulated_adamw_matches_torch_optim(emu):
    n = 4099 - 3                                   # multiple of 4
    p0, g = rnd(n, seed=19), rnd(n, seed=20) * 0.1
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    prm, m, v = p0.clone(), torch.zeros(n), torch.zeros(n)
    for step in (1, 2, 3):
        ref.grad = (g * 0.5).clone()               # grad_scale = 0.5 folds the data-parallel mean
        opt.step()
        assert emu.t2v_adamw_step(prm.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, 0.9, 0.999, 1e-8, 1e-2, step, 0.5, None) == 0
    assert torch.allclose(prm, ref.detach(), rtol=1e-5, atol=1e-7)
    acc = torch.zeros(1)
    assert emu.t2v_sum_squares(g.data_ptr(), n, acc.data_ptr(), None) == 0
    assert abs(acc.item() - (g.double() ** 2).sum().item()) < 1e-4 * acc.item()


def test_emulated_losses_and_row_affine(emu):
    a, b = rnd(2, 4, 3, 5, 8, seed=21), rnd(2, 4, 3, 5, 8, seed=22)
    for name, c in (("mse", None), ("huber", 0.001)):
        loss, grad = torch.zeros(1), torch.empty_like(a)
        if c is None:
            assert emu.t2v_mse_loss_grad(a.data_ptr(), b.data_ptr(), grad.data_ptr(), loss.data_ptr(), a.numel(), 2, 1.0, None) == 0
            ar = a.clone().requires_grad_(True)
            ref = F.mse_loss(ar, b)
        else:
            assert emu.t2v_huber_loss_grad(a.data_ptr(), b.data_ptr(), grad.data_ptr(), loss.data_ptr(), a.numel(), 2, c, 1.0, None) == 0
            ar = a.clone().requires_grad_(True)
            ref = torch.mean(torch.sqrt((ar - b) ** 2 + c ** 2) - c)                        # utils/common_utils.py:302-304
        ref.backward()
        assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item()) and torch.allclose(grad, ar.grad, rtol=1e-4, atol=1e-8), name
    ka, kb = torch.tensor([0.5, -2.0]), torch.tensor([1.5, 0.25])
    out = torch.empty_like(a)
    assert emu.t2v_scale_add_rows(a.data_ptr(), b.data_ptr(), ka.data_ptr(), kb.data_ptr(), out.data_ptr(), 2, a.numel() // 2, 2, None) == 0
    assert torch.allclose(out, a * ka.view(2, 1, 1, 1, 1) + b * kb.view(2, 1, 1, 1, 1), rtol=1e-6, atol=1e-7)


def test_emulated_lcm_step_and_video_post_process(emu):
    """T2VTurboScheduler.step's arithmetic (scheduler/t2v_turbo_scheduler.py:438-460) and app.py:90-94's uint8 conversion."""
    x, e, nz = (rnd(1, 4, 4, 6, 8, seed=s_) for s_ in (59, 60, 61))
    a_t, a_p, c_skip, c_out = torch.tensor(0.0047), torch.tensor(0.35), torch.tensor(2.5e-9), torch.tensor(1.0)
    sb, sa = (1 - a_t).sqrt(), a_t.sqrt()
    den = c_out * ((x - sb * e) / sa) + c_skip * x
    prev = a_p.sqrt() * den + (1 - a_p).sqrt() * nz
    p2, d2 = torch.empty_like(x), torch.empty_like(x)
    assert emu.t2v_lcm_step(x.data_ptr(), e.data_ptr(), nz.data_ptr(), p2.data_ptr(), d2.data_ptr(), x.numel(), 2, float(1.0 / sa), float(sb),
                            float(c_skip), float(c_out), float(a_p.sqrt()), float((1 - a_p).sqrt()), None) == 0
    assert torch.allclose(d2, den, rtol=1e-5, atol=1e-5) and torch.allclose(p2, prev, rtol=1e-5, atol=1e-5)
    v = (rnd(2, 3, 4, 16, 24, seed=43) * 0.8).to(BF16)
    ref = []
    for vid in v:   # app.py:90-94
        t = torch.clamp(vid.float(), -1.0, 1.0).permute(1, 0, 2, 3)
        ref.append((((t + 1.0) / 2.0) * 255).to(torch.uint8).permute(0, 2, 3, 1))
    out = torch.empty(2, 4, 16, 24, 3, dtype=torch.uint8)
    assert emu.t2v_video_to_uint8(v.data_ptr(), 0, out.data_ptr(), 2, 4, 16, 24, None) == 0
    assert torch.equal(out, torch.stack(ref))


def test_emulated_small_cin_conv_and_layout_kernels(emu):
    """The 4-channel latent convolution (dynamic shared memory under emulation), weight packing and the layout conversions."""
    n, h, w, cin, cout = 2, 6, 8, 4, 16
    xw = rnd(n, cin, h, w, seed=23).to(BF16)
    wt = (rnd(cout, cin, 3, 3, seed=24) * 0.2)
    bias = rnd(cout, seed=25)
    x_cl = xw.permute(0, 2, 3, 1).contiguous()
    packed = torch.empty(cout, 9 * cin, dtype=BF16)
    assert emu.t2v_pack_conv_weight(wt.data_ptr(), 2, packed.data_ptr(), cout, cin, 9, None) == 0
    assert torch.equal(packed, wt.reshape(cout, cin, 9).permute(0, 2, 1).reshape(cout, -1).to(BF16))
    out = torch.empty(n, h, w, cout, dtype=BF16)
    assert emu.t2v_conv3x3_small_cin(x_cl.data_ptr(), packed.data_ptr(), bias.data_ptr(), out.data_ptr(), n, h, w, cin, cout, None) == 0
    ref = F.conv2d(xw.float(), wt.to(BF16).float(), bias, padding=1).permute(0, 2, 3, 1)
    close(out, ref, 8e-3, 4e-3, "emulated conv3x3 cin=4")
    # [B, C, T, H, W] <-> frames, channel padding
    lat = rnd(2, 4, 3, h, w, seed=26)
    fr = torch.empty(6, h, w, 64, dtype=BF16)
    assert emu.t2v_bcthw_to_frames_pad(lat.data_ptr(), 2, fr.data_ptr(), 2, 4, 64, 3, h, w, 1.0, None) == 0
    assert torch.equal(fr[..., :4], lat.permute(0, 2, 3, 4, 1).reshape(6, h, w, 4).to(BF16)) and (fr[..., 4:] == 0).all()
    back = torch.empty(2, 4, 3, h, w)
    assert emu.t2v_frames_to_bcthw(fr.data_ptr(), 64, back.data_ptr(), 2, 2, 4, 3, h, w, None) == 0
    assert torch.equal(back, lat.to(BF16).float())
    a_, b_ = rnd(30, 16, seed=27).to(BF16), rnd(30, 8, seed=28).to(BF16)
    cat = torch.empty(30, 24, dtype=BF16)
    assert emu.t2v_concat_channels(a_.data_ptr(), 16, b_.data_ptr(), 8, cat.data_ptr(), 30, None) == 0
    assert torch.equal(cat, torch.cat([a_, b_], 1))
    up = torch.empty(6, 2 * h, 2 * w, 64, dtype=BF16)
    assert emu.t2v_upsample_nearest2x(fr.data_ptr(), up.data_ptr(), 6, h, w, 64, None) == 0
    assert torch.equal(up, fr.repeat_interleave(2, 1).repeat_interleave(2, 2))
    t = torch.tensor([999.0, 519.0, 0.0])
    freqs = torch.exp(-torch.log(torch.tensor(10000.0)) * torch.arange(0, 32, dtype=torch.float32) / 32)
    emb = torch.empty(3, 64)
    assert emu.t2v_sinusoidal_embedding(t.data_ptr(), freqs.data_ptr(), emb.data_ptr(), 3, 32, 0, 0, None) == 0
    arg = t[:, None] * freqs[None]
    assert torch.allclose(emb, torch.cat([torch.cos(arg), torch.sin(arg)], 1), atol=2e-6)      # lvdm/models/utils_diffusion.py:8-32
