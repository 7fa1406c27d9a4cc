"""TEST INFRASTRUCTURE (never imported by the product): a CPU restatement of the CONTRACT of every `t2v_turbo_b200.ops` entry
the training views call (train_unet.StudentUNet, full_train.FullUNet, lora_train, distill, distill_v2), in plain torch.

Purpose: the hand-written backward is ~600 lines of host composition (tapes, layouts, which gradient goes where).  With these
restatements patched over `ops` (tests/test_train_composition_cpu.py) the WHOLE composition runs on CPU in fp32 and is compared
with autograd through the oracle / the reference's gradients at 1e-4 — so a composition bug shows up without a GPU, and the
`-m gpu` tests are left to prove that each kernel meets the contract restated here.  Adjoints are obtained with torch.autograd
on the forward restatement (independent of the formulas the kernels implement).

ACT is the activation dtype the patched modules use (their module-level BF16 is patched to it): float32 for the tight check.
Every function enforces the layout preconditions of the real wrapper (dtype, contiguity, shapes) except `is_cuda`.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

ACT = torch.float32

_TAPS_3X3 = [(kx - 1, ky - 1, 0, 0) for ky in range(3) for kx in range(3)]
_TAPS_T3 = [(0, kt - 1, 0, 0) for kt in range(3)]


def _act(x, name="x", contiguous=True):
    assert x.dtype == ACT, f"{name}: dtype {x.dtype}, expected the activation dtype {ACT}"
    if contiguous:
        assert x.is_contiguous(), f"{name}: must be contiguous"
    return x


def _rows2d(x, name):
    assert x.dtype == ACT and x.dim() == 2 and x.stride(1) == 1, f"{name}: [rows, C] with contiguous channels"
    return x


def _pair(x):
    if isinstance(x, (tuple, list)):
        a, b = x
        return a if b is None else torch.cat([a, b], -1)
    return x


# ----------------------------------------------------------------------------- GEMM layers
def pack_conv_weight(w):
    """[Cout, Cin, *k] -> [Cout, taps * Cin], tap-major K."""
    cout, cin = w.shape[0], w.shape[1]
    return w.reshape(cout, cin, -1).permute(0, 2, 1).reshape(cout, -1).to(ACT).contiguous()


def _unpack(w, cin, taps):
    return w.view(w.shape[0], taps, cin).permute(0, 2, 1)          # [Cout, Cin, taps]


def linear(x, w, bias=None, *, residual=None, **kw):
    assert not kw.get("geglu") and not kw.get("gelu") and kw.get("ln") is None, "fused epilogues are not part of the training path"
    xs = x if isinstance(x, (tuple, list)) else (x, None)
    _act(xs[0])
    x = _pair(x)
    assert w.dtype == ACT and w.is_contiguous() and w.shape[1] == x.shape[1]
    y = x.float() @ w.float().t()
    if bias is not None:
        assert bias.dtype == torch.float32
        y = y + bias
    if residual is not None:
        y = y + residual.reshape(y.shape).float()
    return y.to(ACT)


def conv3x3(x, w, bias=None, *, bias_div=1, residual=None, **kw):
    xs = x if isinstance(x, (tuple, list)) else (x, None)
    _act(xs[0])
    x = _pair(x)
    n, h, wd, c = x.shape
    assert w.dtype == ACT and w.is_contiguous() and w.shape[1] == 9 * c, (w.shape, c)
    if isinstance(xs, tuple) and xs[1] is not None:
        # channel-concatenated pair: per tap the K range is [c0 | c1]
        c0, c1 = xs[0].shape[-1], xs[1].shape[-1]
        wt = w.view(w.shape[0], 9, c0 + c1).permute(0, 2, 1)
    else:
        wt = _unpack(w, c, 9)
    y = F.conv2d(x.permute(0, 3, 1, 2).float(), wt.reshape(w.shape[0], c, 3, 3).float(), padding=1).permute(0, 2, 3, 1)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.dim() == 2 and bias.is_contiguous()
        rows = torch.arange(n) // bias_div
        y = y + bias[rows][:, None, None, :]
    if residual is not None:
        y = y + residual.reshape(y.shape).float()
    return y.to(ACT).contiguous()


def tconv3(x, w, bias=None, *, residual=None, **kw):
    xs = x if isinstance(x, (tuple, list)) else (x, None)
    _act(xs[0])
    x = _pair(x)
    b, t, hw, c = x.shape
    assert w.dtype == ACT and w.is_contiguous() and w.shape[1] == 3 * c
    wt = w.view(w.shape[0], 3, c).permute(0, 2, 1)                 # [Cout, C, 3]  (pairs: per tap [c0 | c1], same view)
    y = F.conv1d(x.permute(0, 2, 3, 1).reshape(b * hw, c, t).float(), wt.float(), padding=1)
    y = y.view(b, hw, -1, t).permute(0, 3, 1, 2)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.dim() == 1
        y = y + bias
    if residual is not None:
        y = y + residual.reshape(y.shape).float()
    return y.to(ACT).contiguous()


def _shift(a, off):
    """a[..points.., C] shifted so that out[p] = a[p + off] with zero fill; off = (x1, x2, x3, x4), x1 the LAST point dim."""
    nd = a.dim() - 1
    out = a
    for j, o in enumerate(off[:nd]):
        if o == 0:
            continue
        dim = nd - 1 - j
        out = torch.roll(out, -o, dim)
        idx = [slice(None)] * out.dim()
        idx[dim] = slice(out.shape[dim] - o, None) if o > 0 else slice(0, -o)
        out = out.clone()
        out[tuple(idx)] = 0
    assert all(o == 0 for o in off[nd:])
    return out


def _wgrad(a, b, out, taps, out_strides, alpha):
    _act(a, "a")
    _act(b, "b")
    assert out.dtype == torch.float32
    assert tuple(a.shape[:-1]) == tuple(b.shape[:-1]) and a.dim() <= 5
    taps = taps if taps is not None else [(0, 0, 0, 0)]
    c, n = a.shape[-1], b.shape[-1]
    js, cs, ts = (int(v) for v in out_strides)
    view = torch.as_strided(out, (n, c, len(taps)), (js, cs, ts if len(taps) > 1 else 1), out.storage_offset())
    b2 = b.reshape(-1, n).float()
    for t, off in enumerate(taps):
        g = b2.t() @ _shift(a, off).reshape(-1, c).float()          # [n, c]
        view[:, :, t] += alpha * g
    return out


def wgrad(a, b, out, *, taps=None, out_strides, alpha=1.0, a_grid=None):
    assert b.shape[-1] <= 64 and b.shape[-1] % 8 == 0 and a.shape[-1] % 8 == 0
    return _wgrad(a, b, out, taps, out_strides, alpha)


def wgrad_wide(a, b, out, *, taps=None, out_strides, alpha=1.0):
    assert b.shape[-1] % 8 == 0 and a.shape[-1] % 8 == 0
    return _wgrad(a, b, out, taps, out_strides, alpha)


# ----------------------------------------------------------------------------- norms
def _gn_f(x, gamma, beta, rows_per_sample, eps, silu, groups):
    rows, c = x.shape
    n = rows // rows_per_sample
    xg = x.float().view(n, rows_per_sample, groups, c // groups)
    mean = xg.mean((1, 3), keepdim=True)
    var = xg.var((1, 3), unbiased=False, keepdim=True)
    y = ((xg - mean) * torch.rsqrt(var + eps)).view(rows, c) * gamma + beta
    return F.silu(y) if silu else y


def groupnorm(x, gamma, beta, *, rows_per_sample, eps, silu, groups=32, **kw):
    x = _pair(x)
    x = x.reshape(-1, x.shape[-1])
    assert x.dtype == ACT and gamma.dtype == torch.float32 and beta.dtype == torch.float32
    assert x.shape[0] % rows_per_sample == 0 and x.shape[1] % groups == 0 and x.shape[1] % 8 == 0
    return _gn_f(x, gamma, beta, rows_per_sample, eps, silu, groups).to(ACT)


def groupnorm_bwd(x, dy, gamma, beta, *, rows_per_sample, eps, silu, groups=32, dx_add=None, out=None, keep_ws=None):
    _rows2d(x, "x"); _rows2d(dy, "dy")
    xr = x.detach().float().requires_grad_(True)
    with torch.enable_grad():
        y = _gn_f(xr, gamma.detach(), beta.detach(), rows_per_sample, eps, silu, groups)
        (dx,) = torch.autograd.grad(y, xr, dy.float())
    if dx_add is not None:
        _rows2d(dx_add, "dx_add")
        dx = dx + dx_add.float()
    if keep_ws is not None:
        keep_ws.append(("gn-stats", x.data_ptr(), rows_per_sample, groups))
    return dx.to(ACT)


def groupnorm_affine_grad(x, dy, gamma, beta, stats_ws, dgamma, dbeta, *, rows_per_sample, eps, silu, groups=32):
    _rows2d(x, "x"); _rows2d(dy, "dy")
    assert stats_ws == ("gn-stats", x.data_ptr(), rows_per_sample, groups), "statistics workspace of a different groupnorm_bwd call"
    assert dgamma.dtype == torch.float32 and dbeta.dtype == torch.float32 and dgamma.is_contiguous() and dbeta.is_contiguous()
    g = gamma.detach().clone().requires_grad_(True)
    b = beta.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        y = _gn_f(x.detach().float(), g, b, rows_per_sample, eps, silu, groups)
        dg, db = torch.autograd.grad(y, (g, b), dy.float())
    dgamma += dg
    dbeta += db


def _ln_f(x, gamma, beta, eps):
    return F.layer_norm(x.float(), (x.shape[-1],), gamma, beta, eps)


def layernorm(x, gamma, beta, eps=1e-5, out=None):
    _rows2d(x, "x")
    assert x.shape[1] in (64, 128, 256, 320, 512, 640, 1024, 1280)
    return _ln_f(x, gamma, beta, eps).to(ACT)


def layernorm_bwd(x, dy, gamma, eps=1e-5, *, dx_add=None, out=None):
    _rows2d(x, "x"); _rows2d(dy, "dy")
    xr = x.detach().float().requires_grad_(True)
    with torch.enable_grad():
        y = _ln_f(xr, gamma.detach(), torch.zeros_like(gamma), eps)
        (dx,) = torch.autograd.grad(y, xr, dy.float())
    if dx_add is not None:
        dx = dx + _rows2d(dx_add, "dx_add").float()
    return dx.to(ACT)


def layernorm_affine_grad(x, dy, dgamma, dbeta, eps=1e-5):
    _rows2d(x, "x"); _rows2d(dy, "dy")
    assert dgamma.dtype == torch.float32 and dbeta.dtype == torch.float32 and dgamma.is_contiguous() and dbeta.is_contiguous()
    xh = F.layer_norm(x.float(), (x.shape[-1],), None, None, eps)
    dgamma += (dy.float() * xh).sum(0)
    dbeta += dy.float().sum(0)


# ----------------------------------------------------------------------------- attention
def _attn_f(q, k, v, heads, scale):
    bq, lq, inner = q.shape
    bk, lk, _ = k.shape
    qh = q.float().view(bq, lq, heads, 64).transpose(1, 2)
    kh = k.float().view(bk, lk, heads, 64).transpose(1, 2)
    vh = v.float().view(bk, lk, heads, 64).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) * scale
    p = s.softmax(-1)
    return (p @ vh).transpose(1, 2).reshape(bq, lq, inner), s


def attention(q, k, v, *, heads, scale, kv_batch_div=1, out=None, causal=False, lse2=None):
    assert kv_batch_div == 1 and not causal
    for t in (q, k, v):
        assert t.dtype == ACT and t.stride(2) == 1
    assert q.shape[2] == heads * 64 and q.shape[0] == k.shape[0]
    o, s = _attn_f(q, k, v, heads, scale)
    if lse2 is not None:
        assert lse2.dtype == torch.float32 and tuple(lse2.shape) == (q.shape[0], heads, q.shape[1])
        lse2.copy_(torch.logsumexp(s, -1) / math.log(2.0))
    return o.to(ACT)


def attention_bwd(q, k, v, o, d_o, lse2, *, heads, scale, kv_batch_div=1, need_dq=True, need_dkv=True):
    assert kv_batch_div == 1
    qr, kr, vr = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    with torch.enable_grad():
        y, s = _attn_f(qr, kr, vr, heads, scale)
        dq, dk, dv = torch.autograd.grad(y, (qr, kr, vr), d_o.float().reshape(y.shape))
    assert torch.allclose(lse2, (torch.logsumexp(s, -1) / math.log(2.0)).detach(), atol=1e-4), "lse2 is not this forward's"
    assert torch.allclose(o.float(), y.detach(), atol=1e-4), "o is not this forward's output"
    return dq.to(ACT).contiguous(), dk.to(ACT).contiguous(), dv.to(ACT).contiguous()


def _tattn_f(q, k, v, b, t, hw, heads, scale):
    def seq(x):     # [(b t hw), H*64] -> [(b hw), t, H*64]
        return x.float().view(b, t, hw, -1).permute(0, 2, 1, 3).reshape(b * hw, t, -1)
    o, _ = _attn_f(seq(q), seq(k), seq(v), heads, scale)
    return o.view(b, hw, t, -1).permute(0, 2, 1, 3).reshape(b * t * hw, -1)


def _tprobs_f(q, k, b, t, hw, heads, scale):
    """softmax(scale q k^T) in the export's layout [(b hw heads), t, t] (the reference's "(b h) i j" with b = pixels)."""
    def seq(x):
        return x.float().view(b, t, hw, heads, 64).permute(0, 2, 3, 1, 4).reshape(b * hw * heads, t, 64)
    return torch.softmax(seq(q) @ seq(k).transpose(1, 2) * scale, -1)


def attention_temporal(q, k, v, *, b, t, hw, heads, scale, out=None, probs=None):
    for x in (q, k, v):
        assert x.dtype == ACT and x.dim() == 2 and x.stride(1) == 1 and x.shape == (b * t * hw, heads * 64)
    if probs is not None:
        assert probs.is_contiguous() and tuple(probs.shape) == (b * hw * heads, t, t)
        probs.copy_(_tprobs_f(q, k, b, t, hw, heads, scale).to(probs.dtype))
    return _tattn_f(q, k, v, b, t, hw, heads, scale).to(ACT)


def attention_temporal_probs_bwd(q, k, d_probs, *, b, t, hw, heads, scale):
    assert q.dtype == ACT and k.dtype == ACT and q.stride(1) == 1 and k.stride(1) == 1 and t <= 16
    assert d_probs.dtype == torch.float32 and d_probs.is_contiguous() and tuple(d_probs.shape) == (b * hw * heads, t, t)
    qr, kr = q.detach().float().requires_grad_(True), k.detach().float().requires_grad_(True)
    with torch.enable_grad():
        dq, dk = torch.autograd.grad(_tprobs_f(qr, kr, b, t, hw, heads, scale), (qr, kr), d_probs)
    return dq.to(ACT).contiguous(), dk.to(ACT).contiguous()


def attention_temporal_bwd(q, k, v, d_o, *, b, t, hw, heads, scale):
    qr, kr, vr = (x.detach().float().requires_grad_(True) for x in (q, k, v))
    with torch.enable_grad():
        y = _tattn_f(qr, kr, vr, b, t, hw, heads, scale)
        dq, dk, dv = torch.autograd.grad(y, (qr, kr, vr), d_o.float().reshape(y.shape))
    return dq.to(ACT).contiguous(), dk.to(ACT).contiguous(), dv.to(ACT).contiguous()


def bmm_nt(a, b, *, out=None, alpha=1.0, block_n=0):
    assert a.dtype == ACT and b.dtype == ACT and a.stride(2) == 1 and b.stride(2) == 1
    assert a.shape[0] == b.shape[0] and a.shape[2] == b.shape[2] and a.shape[2] % 64 == 0, (a.shape, b.shape)
    return (alpha * (a.float() @ b.float().transpose(1, 2))).to(ACT)


def softmax_rows_(x, scale=1.0):
    assert x.dtype == ACT and x.is_contiguous()
    x.copy_(torch.softmax(x.float() * scale, -1).to(ACT))
    return x


def softmax_bwd_rows_(dp, p, scale=1.0):
    assert dp.dtype == ACT and p.dtype == ACT and dp.is_contiguous() and p.is_contiguous() and dp.shape == p.shape
    pf, df = p.float(), dp.float()
    dp.copy_((scale * pf * (df - (df * pf).sum(-1, keepdim=True))).to(ACT))
    return dp


def conv3x3_small_cin(x, w, bias, cout):
    _act(x)
    n, h, wd, cin = x.shape
    assert w.dtype == ACT and tuple(w.shape) == (cout, 9 * cin) and cin in (4, 8)
    y = F.conv2d(x.permute(0, 3, 1, 2).float(), _unpack(w, cin, 9).reshape(cout, cin, 3, 3).float(), bias, padding=1)
    return y.permute(0, 2, 3, 1).to(ACT).contiguous()


def bcthw_to_frames_mix(z, scale, mix, bias):
    b, c, t, hh, ww = z.shape
    assert mix.dtype == torch.float32 and bias.dtype == torch.float32 and tuple(mix.shape) == (bias.numel(), c)
    fr = (z.float() * scale).permute(0, 2, 3, 4, 1).reshape(b * t, hh, ww, c)
    return (fr @ mix.t() + bias).to(ACT).contiguous()


# ----------------------------------------------------------------------------- elementwise / layout
def _geglu_f(pre):
    i = pre.shape[1] // 2
    return pre[:, :i] * F.gelu(pre[:, i:])


def geglu(pre, dout=None, out=None):
    _rows2d(pre, "pre")
    if dout is None:
        return _geglu_f(pre.float()).to(ACT)
    _rows2d(dout, "dout")
    pr = pre.detach().float().requires_grad_(True)
    with torch.enable_grad():
        (d,) = torch.autograd.grad(_geglu_f(pr), pr, dout.float())
    return d.to(ACT)


def add(a, b, out=None):
    _rows2d(a, "a"); _rows2d(b, "b")
    assert a.shape == b.shape
    return (a.float() + b.float()).to(ACT)


def silu(a, out=None):
    _rows2d(a, "a")
    return F.silu(a.float()).to(ACT)


def silu_bwd(pre, dy, out=None):
    _rows2d(pre, "pre"); _rows2d(dy, "dy")
    p = pre.float()
    sg = torch.sigmoid(p)
    return (dy.float() * sg * (1 + p * (1 - sg))).to(ACT)


def colsum_samples(x, rows_per_sample, out=None):
    _rows2d(x, "x")
    rows, c = x.shape
    assert rows % rows_per_sample == 0 and c % 8 == 0
    s = x.float().view(rows // rows_per_sample, rows_per_sample, c).sum(1)
    if out is None:
        return s
    assert out.dtype == torch.float32 and tuple(out.shape) == tuple(s.shape)
    out += s
    return out


def resample2x(x, mode):
    _act(x)
    assert x.dim() == 4
    n, h, w, c = x.shape
    if mode == "sub":
        return x[:, ::2, ::2].contiguous()
    if mode == "stuff":
        out = x.new_zeros((n, 2 * h, 2 * w, c))
        out[:, ::2, ::2] = x
        return out
    if mode == "pool":
        return x.float().view(n, h // 2, 2, w // 2, 2, c).sum((2, 4)).to(ACT)
    raise ValueError(mode)


def upsample_nearest2x(x):
    _act(x)
    return x.repeat_interleave(2, 1).repeat_interleave(2, 2).contiguous()


def concat_channels(a, b):
    assert a.dtype == ACT and b.dtype == ACT and a.shape[:-1] == b.shape[:-1]
    return torch.cat([a, b], -1).contiguous()


def bcthw_to_frames_pad(x, c_pad, scale=1.0):
    b, c, t, h, w = x.shape
    out = torch.zeros((b * t, h, w, c_pad), dtype=ACT)
    out[..., :c] = (x.float() * scale).permute(0, 2, 3, 4, 1).reshape(b * t, h, w, c).to(ACT)
    return out


def frames_to_bcthw(x, b, c, dtype):
    n, h, w, cp = x.shape
    t = n // b
    return x[..., :c].reshape(b, t, h, w, c).permute(0, 4, 1, 2, 3).to(dtype).contiguous()


def sinusoidal_embedding(t, freqs, *, sin_first=False, round_bf16=True):
    assert t.dtype == torch.float32 and freqs.dtype == torch.float32
    arg = t[:, None] * freqs[None, :]
    s, co = torch.sin(arg), torch.cos(arg)
    if round_bf16 and ACT == torch.bfloat16:
        s, co = s.bfloat16().float(), co.bfloat16().float()
    return torch.cat([s, co] if sin_first else [co, s], 1)


def scale_mask(x, scale, mask=None, out=None):
    _act(x)
    y = x.float() * scale
    if mask is not None:
        assert mask.dtype == torch.uint8 and mask.numel() == x.numel()
        y = y * mask.view(x.shape)
    return y.to(ACT)


def dropout_advance(device):
    pass


def dropout_scale(x, p, scale=1.0, out=None, addend=None):
    _act(x)
    keep = torch.empty(x.shape, dtype=torch.uint8).bernoulli_(1.0 - p)
    y = x.float() * (scale / (1.0 - p)) * keep
    if addend is not None:
        y = y + addend.float().view(x.shape)
    return y.to(ACT), keep


def scale_add_rows(x, a, y=None, b=None):
    assert x.is_contiguous() and a.dtype == torch.float32 and a.numel() == x.shape[0]
    sh = (-1,) + (1,) * (x.dim() - 1)
    out = x.float() * a.view(sh)
    if y is not None:
        assert y.shape == x.shape and y.dtype == x.dtype and b.numel() == x.shape[0]
        out = out + y.float() * b.view(sh)
    return out.to(x.dtype)


# ----------------------------------------------------------------------------- losses / optimizer
def mse_loss_grad(a, b, *, want_grad=True, grad_scale=1.0):
    assert a.shape == b.shape and a.dtype == b.dtype
    d = a.float() - b.float()
    return (d * d).mean().view(1), ((2.0 * d / d.numel()) * grad_scale).to(a.dtype) if want_grad else None


def huber_loss_grad(a, b, huber_c=0.001, *, want_grad=True, grad_scale=1.0):
    assert a.shape == b.shape and a.dtype == b.dtype
    d = a.float() - b.float()
    r = torch.sqrt(d * d + huber_c ** 2)
    return (r - huber_c).mean().view(1), ((d / r / d.numel()) * grad_scale).to(a.dtype) if want_grad else None


def sum_squares(x, out=None):
    assert x.dtype == torch.float32 and x.is_contiguous()
    s = (x.double() ** 2).sum().float().view(1)
    if out is None:
        return s
    out += s
    return out


def adamw_step(param, grad, exp_avg, exp_avg_sq, *, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, step, grad_scale=1.0):
    for t in (param, grad, exp_avg, exp_avg_sq):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == param.numel()
        assert t.data_ptr() % 16 == 0, "the fused AdamW reads float4: slices must start on a 16-byte boundary"
    g = grad * grad_scale
    param.mul_(1.0 - lr * weight_decay)
    exp_avg.mul_(betas[0]).add_(g, alpha=1.0 - betas[0])
    exp_avg_sq.mul_(betas[1]).addcmul_(g, g, value=1.0 - betas[1])
    bc1, bc2 = 1.0 - betas[0] ** step, 1.0 - betas[1] ** step
    denom = (exp_avg_sq.sqrt() / math.sqrt(bc2)).add_(eps)
    param.addcdiv_(exp_avg, denom, value=-lr / bc1)


def ema_update(target, src, rate):
    assert target.dtype == torch.float32 and src.dtype == torch.float32 and target.numel() == src.numel()
    target.mul_(rate).add_(src, alpha=1.0 - rate)
    return target


ALL = [n for n, f in list(globals().items()) if callable(f) and not n.startswith("_") and n not in ("F",)]


def install(monkeypatch, act=torch.float32):
    """Patch the restatements over t2v_turbo_b200.ops and the activation dtype over the training modules."""
    global ACT
    ACT = act
    from t2v_turbo_b200 import full_train, lora_train, ops, train_unet
    for name in ALL:
        if name in ("install",):
            continue
        monkeypatch.setattr(ops, name, globals()[name], raising=True)
    from t2v_turbo_b200 import motion_prior, vae_train
    for mod in (train_unet, full_train, lora_train, vae_train, motion_prior):
        monkeypatch.setattr(mod, "BF16", act)
    for mod in (train_unet, full_train, vae_train, motion_prior):
        monkeypatch.setattr(mod, "_require_cuda", lambda what, device: None)
    return ops
