"""The SIMT kernels of csrc/train_bwd.cu and csrc/train_full.cu EXECUTED on the CPU: the two sources are compiled unchanged by
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
         "t2v_groupnorm_affine_grad", "t2v_layernorm_affine_grad", "t2v_ema_update")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    from t2v_turbo_b200 import _lib
    out = tmp_path_factory.mktemp("cuda_emu") / "libt2v_emu.so"
    csrc = os.path.join(ROOT, "t2v_turbo_b200", "csrc")
    cmd = ["g++", "-std=c++20", "-O1", "-x", "c++", "-DT2V_HOST_EMU", "-I", os.path.join(ROOT, "tests", "cuda_emu"), "-shared", "-fPIC",
           "-pthread", os.path.join(csrc, "train_bwd.cu"), os.path.join(csrc, "train_full.cu"), "-o", str(out)]
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


@pytest.mark.parametrize("n,offset", [(4096, 0), (100003, 0), (4099, 1), (3, 0), (1, 0)])
def test_ema_update_kernel_under_emulation(emu, n, offset):
    tgt, src = rnd(n + offset, seed=16)[offset:], rnd(n + offset, seed=17)[offset:]     # offset 1: unaligned for float4 -> scalar path
    ref = tgt.clone().mul_(0.95).add_(src, alpha=0.05)
    assert emu.t2v_ema_update(tgt.data_ptr(), src.data_ptr(), n, 0.95, None) == 0
    assert torch.allclose(tgt, ref, rtol=1e-6, atol=1e-7)
    assert emu.t2v_ema_update(tgt.data_ptr(), src.data_ptr(), n, 1.5, None) < 0
