"""CPU checks of the host-side algebra that the GPU kernels rely on (no GPU, no shared library calls):
tile planning, the phase decomposition of upsample + conv, and the LayerNorm fold."""
import itertools
import math

import pytest
import torch
import torch.nn.functional as F

from t2v_turbo_b200 import ops


@pytest.mark.parametrize("sizes", [(64, 40, 16, 1), (32, 20, 16, 1), (16, 10, 16, 1), (8, 5, 16, 1), (2560, 16, 1, 1),
                                   (40, 16, 1, 1), (512, 320, 16, 1), (4, 4, 4, 1), (160, 16, 1, 1)])
def test_plan_box_covers_grid_with_at_most_128_rows(sizes):
    box = ops.plan_box(sizes)
    rows = math.prod(box)
    assert rows <= 128 and rows % 8 == 0
    tiles = math.prod(-(-s // b) for s, b in zip(sizes, box))
    assert tiles * rows >= math.prod(sizes)
    # the UNet / VAE geometries tile without waste
    if sizes in [(64, 40, 16, 1), (32, 20, 16, 1), (16, 10, 16, 1), (8, 5, 16, 1), (2560, 16, 1, 1), (512, 320, 16, 1)]:
        assert tiles * 128 == math.prod(sizes)


@pytest.mark.parametrize("n,h,w,cin,cout", [(1, 5, 7, 4, 6), (2, 8, 8, 3, 5)])
def test_upsample_conv_equals_four_presummed_phase_convs(n, h, w, cin, cout):
    """ops.pack_upconv_weight / ops.upconv3x3: nearest-2x upsample + 3x3 conv (pad 1) == four 2x2 convs on the
    low-resolution input, one per output parity, zero padding -> out-of-range source pixels (openaimodel3d.py:96-108)."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), wt, padding=1)
    phases = ops.pack_upconv_weight(wt).float()                      # [4][cout][4*cin], tap-major, bf16-rounded
    out = torch.zeros_like(ref)
    xp = F.pad(x, (1, 1, 1, 1))                                       # source pixels -1 .. h (zero outside)
    for py, px in itertools.product(range(2), range(2)):
        taps = [(dy, dx) for dy, _ in ops._UP_ROWS[py] for dx, _ in ops._UP_ROWS[px]]
        wp = phases[2 * py + px].view(cout, 4, cin)
        acc = torch.zeros(n, cout, h, w)
        for t, (dy, dx) in enumerate(taps):
            src = xp[:, :, 1 + dy:1 + dy + h, 1 + dx:1 + dx + w]
            acc += torch.einsum("oc,nchw->nohw", wp[:, t], src)
        out[:, :, py::2, px::2] = acc
    assert (out - ref).abs().max().item() < 3e-2 * ref.abs().max().item()        # bf16 rounding of the summed weights


def test_layernorm_fold_identity():
    """ops.fold_layernorm: LN(x) W^T + b == rstd (x W'^T) + (-rstd mean) colsum + b'  (attention.py:279-281)."""
    g = torch.Generator().manual_seed(1)
    m, c, n = 37, 64, 48
    x = torch.randn(m, c, generator=g) * 1.5 + 0.7
    w = torch.randn(n, c, generator=g) / 8
    b = torch.randn(n, generator=g)
    gamma, beta = 1 + 0.2 * torch.randn(c, generator=g), 0.1 * torch.randn(c, generator=g)
    wp, bp, cs = ops.fold_layernorm(w, b, gamma, beta)
    mean = x.mean(1, keepdim=True)
    rstd = (x.var(1, unbiased=False, keepdim=True) + 1e-5).rsqrt()
    ours = rstd * (x @ wp.float().t()) + (-rstd * mean) * cs[None, :] + bp[None, :]
    ref = F.linear(F.layer_norm(x, (c,), gamma, beta, 1e-5), w, b)
    assert (ours - ref).abs().max().item() < 2e-2 * ref.abs().max().item()         # bf16 rounding of W * gamma only
    # with unrounded folded weights the identity is exact to fp32
    w_exact = w * gamma[None, :]
    exact = rstd * (x @ w_exact.t()) + (-rstd * mean) * w_exact.sum(1)[None, :] + (w @ beta + b)[None, :]
    torch.testing.assert_close(exact, ref, rtol=1e-4, atol=1e-4)


def test_gn_fuse_policy_defaults_off():
    assert ops.GN_FUSE in ("off", "conv", "all")
    if ops.GN_FUSE == "off":
        assert not ops.gn_fuse_producer(1 << 20, (64, 40, 16, 1))


def test_unet_statistics_workspace_per_geometry():
    """UNetModel keeps one statistics workspace per input geometry (captured CUDA graphs hold raw pointers into it):
    sizing pass -> one buffer -> reuse with a single zeroing; another geometry gets its own buffer."""
    from oracle.configs import UNET_CONFIGS
    from t2v_turbo_b200.unet import UNetModel
    m = UNetModel(**UNET_CONFIGS["small"]["cfg"])
    dev = torch.device("cpu")

    def fwd(key, rows):
        m._ln_begin(key)
        outs = [m._ln_slice(r, dev) for r in rows]
        for o in outs:
            o += 1.0
        return outs

    ka, kb = (1, 4, 8, 8, dev), (2, 4, 8, 8, dev)
    fwd(ka, [10, 20, 30])
    assert m._ln_states[ka]["buf"] is None and m._ln_states[ka]["need"] == 120
    a2 = fwd(ka, [10, 20, 30])
    buf_a = m._ln_states[ka]["buf"]
    assert buf_a.numel() == 120 and float(buf_a.sum()) == 120.0
    assert a2[1].data_ptr() == buf_a.data_ptr() + 20 * 4
    fwd(kb, [16])
    fwd(kb, [16])
    fwd(ka, [10, 20, 30])
    assert m._ln_states[ka]["buf"] is buf_a and float(buf_a.sum()) == 120.0
    assert m._ln_states[kb]["buf"] is not None and m._ln_states[kb]["buf"] is not buf_a
