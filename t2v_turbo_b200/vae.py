"""Drop-in KL-VAE on B200: `AutoencoderKL.decode` / `decode_first_stage_2DAE` and `AutoencoderKL.encode`.

Mirrors lvdm/models/autoencoder.py:110-113 and lvdm/modules/networks/ae_modules.py:506-641 (Decoder,
ResnetBlock :146-203, AttnBlock :29-73, Upsample :108-122): same constructor config (`ddconfig`,
`embed_dim`), same state-dict keys (`post_quant_conv.*`, `decoder.conv_in.*`, `decoder.mid.block_1.*`,
`decoder.up.{i}.block.{j}.*`, `decoder.up.{i}.upsample.conv.*`, ...).  The nn modules are parameter
containers; the arithmetic is libt2v_b200.so kernels over channels-last bf16 with ALL frames batched
(the reference decodes frame by frame in a Python loop, ddpm3d.py:671-677).
The encode side (training only: autoencoder.py:103-108, Encoder ae_modules.py:381-503, DiagonalGaussian
distributions.py:24-42) uses the same kernels: `encoder.*` / `quant_conv.*` keys, Downsample as a stride-2 conv over a
parity view with the reference's right/bottom zero padding, conv_out and quant_conv folded into one GEMM with fp32
moments, posterior sample / mode in one elementwise kernel.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from ._lib import lib, stream_ptr, DTYPE_CODE

BF16 = torch.bfloat16


def _norm(ch):
    return nn.GroupNorm(32, ch, eps=1e-6, affine=True)


class ResnetBlock(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = _norm(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.norm2 = _norm(out_channels)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        if in_channels != out_channels:
            self.nin_shortcut = nn.Conv2d(in_channels, out_channels, 1)


class AttnBlock(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.in_channels = ch
        self.norm = _norm(ch)
        self.q, self.k, self.v, self.proj_out = (nn.Conv2d(ch, ch, 1) for _ in range(4))


class UpsampleConv(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)


class DownsampleConv(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=0)


class Encoder(nn.Module):
    """Parameter container with the reference's key layout (ae_modules.py:381-468)."""

    def __init__(self, *, ch, out_ch=None, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions=(), dropout=0.0,
                 resamp_with_conv=True, in_channels=3, resolution=256, z_channels, double_z=True, **ignored):
        super().__init__()
        if len(attn_resolutions) or not resamp_with_conv:
            raise NotImplementedError("Encoder(B200): attn_resolutions / resamp_with_conv=False are not used by VC2")
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        self.conv_in = nn.Conv2d(in_channels, ch, 3, padding=1)
        in_ch_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        block_in = ch
        for i_level in range(self.num_resolutions):
            block_in = ch * in_ch_mult[i_level]
            block_out = ch * ch_mult[i_level]
            down = nn.Module()
            down.block = nn.ModuleList()
            down.attn = nn.ModuleList()
            for _ in range(num_res_blocks):
                down.block.append(ResnetBlock(block_in, block_out))
                block_in = block_out
            if i_level != self.num_resolutions - 1:
                down.downsample = DownsampleConv(block_in)
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(block_in, block_in)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(block_in, block_in)
        self.norm_out = _norm(block_in)
        self.conv_out = nn.Conv2d(block_in, 2 * z_channels if double_z else z_channels, 3, padding=1)


class Decoder(nn.Module):
    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions=(), dropout=0.0,
                 resamp_with_conv=True, in_channels=3, resolution=256, z_channels, **ignored):
        super().__init__()
        if len(attn_resolutions) or not resamp_with_conv:
            raise NotImplementedError("Decoder(B200): attn_resolutions / resamp_with_conv=False are not used by VC2")
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        block_in = ch * ch_mult[-1]
        self.conv_in = nn.Conv2d(z_channels, block_in, 3, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(block_in, block_in)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(block_in, block_in)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block_out = ch * ch_mult[i_level]
            up = nn.Module()
            up.block = nn.ModuleList()
            up.attn = nn.ModuleList()
            for _ in range(num_res_blocks + 1):
                up.block.append(ResnetBlock(block_in, block_out))
                block_in = block_out
            if i_level != 0:
                up.upsample = UpsampleConv(block_in)
            self.up.insert(0, up)
        self.norm_out = _norm(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, 3, padding=1)


def _f32(t):
    return t.detach().float().contiguous()


class _StatsPool:
    """Zeroed fp32 workspace for the per-frame channel sums the GEMM epilogues accumulate (one memset per call;
    allocated on first use — nothing when the fusion is off, see ops.GN_FUSE)."""

    def __init__(self, device, nfloats):
        self.device, self.nfloats, self.buf, self.cur = device, nfloats, None, 0

    def take(self, frames, channels, k_total, grid, sdims=(2,)):
        if not ops.gn_fuse_producer(k_total, grid, sample_dims=sdims):
            return None
        if self.buf is None:
            self.buf = torch.zeros((self.nfloats,), device=self.device, dtype=torch.float32)
        n = frames * channels * 2
        if self.cur + n > self.buf.numel():
            raise RuntimeError("AutoencoderKL(B200): statistics workspace exhausted")
        v = self.buf[self.cur:self.cur + n].view(frames, channels, 2)
        self.cur += (n + 3) // 4 * 4
        return v


def _w2d(w):
    return w.detach().reshape(w.shape[0], -1).to(BF16).contiguous()


class _PRes:
    def __init__(self, rb: ResnetBlock):
        self.cin, self.cout = rb.in_channels, rb.out_channels
        self.gn1 = (_f32(rb.norm1.weight), _f32(rb.norm1.bias))
        self.w1, self.b1 = ops.pack_conv_weight(rb.conv1.weight.detach()), _f32(rb.conv1.bias).view(1, -1)
        self.gn2 = (_f32(rb.norm2.weight), _f32(rb.norm2.bias))
        self.w2, self.b2 = ops.pack_conv_weight(rb.conv2.weight.detach()), _f32(rb.conv2.bias).view(1, -1)
        self.w_nin = _w2d(rb.nin_shortcut.weight) if self.cin != self.cout else None
        self.b_nin = _f32(rb.nin_shortcut.bias) if self.cin != self.cout else None


class DiagonalGaussianDistribution:
    """The reference's posterior object (lvdm/distributions.py:24-72) over the fp32 channels-last moments
    [N, h, w, 2*zc] the encoder GEMM wrote.  `sample` / `mode` are one elementwise kernel (t2v_gaussian_sample) and return
    NCHW in the dtype of the encoded frames; the attribute views (`mean`, `logvar`, `std`, `var`, `parameters`) and
    `kl` / `nll` (loss-side helpers outside the hot path) are plain torch on the moments."""

    def __init__(self, moments_nhwc, dtype, deterministic=False):
        self._m, self._dtype, self.deterministic = moments_nhwc, dtype, deterministic
        self._zc = moments_nhwc.shape[-1] // 2

    def _out(self, noise):
        n, h, w, _ = self._m.shape
        return ops.gaussian_sample(self._m, noise, b=n, t=1, zc=self._zc, scale=1.0, dtype=self._dtype).squeeze(2)

    def sample(self, noise=None):
        if self.deterministic:
            return self.mode()
        if noise is None:   # the reference draws on the CPU and moves the tensor (distributions.py:38-41)
            n, h, w, _ = self._m.shape
            noise = torch.randn((n, self._zc, h, w))
        return self._out(noise)

    def mode(self):
        return self._out(None)

    @property
    def parameters(self):
        return self._m.permute(0, 3, 1, 2).to(self._dtype)

    @property
    def mean(self):
        return self._m[..., :self._zc].permute(0, 3, 1, 2).to(self._dtype)

    @property
    def logvar(self):
        return torch.clamp(self._m[..., self._zc:], -30.0, 20.0).permute(0, 3, 1, 2).to(self._dtype)

    @property
    def std(self):
        return torch.zeros_like(self.mean) if self.deterministic else torch.exp(0.5 * self.logvar)

    @property
    def var(self):
        return torch.zeros_like(self.mean) if self.deterministic else torch.exp(self.logvar)

    def kl(self, other=None):
        if self.deterministic:
            return torch.Tensor([0.0])
        if other is None:
            return 0.5 * torch.sum(torch.pow(self.mean, 2) + self.var - 1.0 - self.logvar, dim=[1, 2, 3])
        return 0.5 * torch.sum(torch.pow(self.mean - other.mean, 2) / other.var + self.var / other.var - 1.0 - self.logvar
                               + other.logvar, dim=[1, 2, 3])

    def nll(self, sample, dims=(1, 2, 3)):
        if self.deterministic:
            return torch.Tensor([0.0])
        import math
        return 0.5 * torch.sum(math.log(2.0 * math.pi) + self.logvar + torch.pow(sample - self.mean, 2) / self.var, dim=list(dims))


class AutoencoderKL(nn.Module):
    """AutoencoderKL (reference ctor: ddconfig, lossconfig, embed_dim; lossconfig ignored): decode and encode."""

    def __init__(self, ddconfig, embed_dim, lossconfig=None, **ignored):
        super().__init__()
        self.ddconfig, self.embed_dim = dict(ddconfig), embed_dim
        self.encoder = Encoder(**ddconfig)
        self.decoder = Decoder(**ddconfig)
        self.quant_conv = nn.Conv2d(2 * ddconfig["z_channels"], 2 * embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        self._packed = None
        self._packed_enc = None
        self.weight_generation = 0   # bumped whenever the packed weights are dropped: CUDA graphs key on it

    def invalidate_packed(self):
        self._packed = self._packed_enc = None
        self.weight_generation = getattr(self, "weight_generation", 0) + 1

    def _apply(self, fn, *a, **k):
        self.invalidate_packed()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self.invalidate_packed()
        return super().load_state_dict(*a, **k)

    @staticmethod
    def _pack_attn(a):
        c = a.in_channels
        wo = a.proj_out.weight.detach().float().reshape(c, c)
        # v bias folded through the softmax (rows sum to 1) into the output projection bias
        return dict(gn=(_f32(a.norm.weight), _f32(a.norm.bias)),
                    w_qk=torch.cat([_w2d(a.q.weight), _w2d(a.k.weight)], 0).contiguous(),
                    b_qk=torch.cat([_f32(a.q.bias), _f32(a.k.bias)], 0).contiguous(),
                    w_v=_w2d(a.v.weight), w_o=_w2d(a.proj_out.weight),
                    b_o=(_f32(a.proj_out.bias) + wo @ _f32(a.v.bias)).contiguous(), c=c)

    @torch.no_grad()
    def pack_encoder(self):
        e = self.encoder
        dev = e.conv_in.weight.device
        if dev.type != "cuda":
            raise RuntimeError("AutoencoderKL(B200) runs on a CUDA device only (no CPU fallback)")
        P = {}
        cin = e.conv_in.weight.shape[1]
        if cin > 4:
            raise NotImplementedError("Encoder(B200): more than 4 input channels")
        # RGB frames are padded to 4 channels (zero weights for the 4th) so the direct small-Cin conv kernel applies
        w_in = torch.zeros(e.conv_in.weight.shape[0], 4, 3, 3, device=dev, dtype=torch.float32)
        w_in[:, :cin] = e.conv_in.weight.detach().float()
        P["conv_in"] = (ops.pack_conv_weight(w_in), _f32(e.conv_in.bias), w_in.shape[0])
        P["down"] = []
        for i_level in range(e.num_resolutions):
            dn = e.down[i_level]
            blocks = [_PRes(b) for b in dn.block]
            ds = None
            if i_level != e.num_resolutions - 1:
                ds = (ops.pack_conv_weight(dn.downsample.conv.weight.detach()), _f32(dn.downsample.conv.bias))
            P["down"].append((blocks, ds))
        P["mid1"], P["mid2"] = _PRes(e.mid.block_1), _PRes(e.mid.block_2)
        P["attn"] = self._pack_attn(e.mid.attn_1)
        P["norm_out"] = (_f32(e.norm_out.weight), _f32(e.norm_out.bias))
        # quant_conv (1x1) composed with conv_out: moments = (Wq Wc) * x + (Wq bc + bq)
        wq = self.quant_conv.weight.detach().float().reshape(self.quant_conv.weight.shape[0], -1)
        wc = e.conv_out.weight.detach().float()
        w_fold = torch.einsum("oc,cikl->oikl", wq, wc)
        b_fold = wq @ e.conv_out.bias.detach().float() + self.quant_conv.bias.detach().float()
        P["conv_out"] = (ops.pack_conv_weight(w_fold), b_fold.view(1, -1).contiguous(), w_fold.shape[0])
        self._packed_enc = P
        return self

    @torch.no_grad()
    def pack(self):
        d = self.decoder
        dev = d.conv_in.weight.device
        if dev.type != "cuda":
            raise RuntimeError("AutoencoderKL(B200) runs on a CUDA device only (no CPU fallback)")
        P = {}
        zc = self.post_quant_conv.weight.shape[0]
        P["pq_w"] = _f32(self.post_quant_conv.weight).reshape(zc, -1)
        P["pq_b"] = _f32(self.post_quant_conv.bias)
        P["conv_in"] = (ops.pack_conv_weight(d.conv_in.weight.detach()), _f32(d.conv_in.bias), d.conv_in.weight.shape[0])
        P["mid1"], P["mid2"] = _PRes(d.mid.block_1), _PRes(d.mid.block_2)
        P["attn"] = self._pack_attn(d.mid.attn_1)
        P["up"] = []
        for i_level in range(d.num_resolutions):
            up = d.up[i_level]
            blocks = [_PRes(b) for b in up.block]
            ups = None
            if i_level != 0:
                ups = (ops.pack_upconv_weight(up.upsample.conv.weight), _f32(up.upsample.conv.bias))
            P["up"].append((blocks, ups))
        P["norm_out"] = (_f32(d.norm_out.weight), _f32(d.norm_out.bias))
        P["conv_out"] = (ops.pack_conv_weight(d.conv_out.weight.detach()), _f32(d.conv_out.bias).view(1, -1), d.conv_out.weight.shape[0])
        self._packed = P
        return self

    # ------------------------------------------------------------------ pieces
    # Every conv / projection accumulates the per-frame channel sums of its output in its epilogue (`stats`), so each
    # GroupNorm ("Normalize", ae_modules.py:16-19) runs only its apply pass.  `pool` hands out zeroed fp32 slices.
    @staticmethod
    def _res(pr: _PRes, h, hs, pool):
        n, hh, ww, c = h.shape
        hw = hh * ww
        t = ops.groupnorm(h.view(-1, c), pr.gn1[0], pr.gn1[1], rows_per_sample=hw, eps=1e-6, silu=True, chan_sums=hs)
        s1 = pool.take(n, pr.cout, 9 * c, (ww, hh, n, 1))
        t = ops.conv3x3(t.view(n, hh, ww, c), pr.w1, pr.b1, bias_div=n, stats=s1)
        t = ops.groupnorm(t.view(-1, pr.cout), pr.gn2[0], pr.gn2[1], rows_per_sample=hw, eps=1e-6, silu=True, chan_sums=s1)
        res = h if pr.w_nin is None else ops.linear(h.view(-1, c), pr.w_nin, pr.b_nin).view(n, hh, ww, pr.cout)
        s2 = pool.take(n, pr.cout, 9 * pr.cout, (ww, hh, n, 1))
        return ops.conv3x3(t.view(n, hh, ww, pr.cout), pr.w2, pr.b2, bias_div=n, residual=res, stats=s2), s2

    @staticmethod
    def _attn(pa, h, hs, pool):
        n, hh, ww, c = h.shape
        hw = hh * ww
        x = h.view(-1, c)
        xn = ops.groupnorm(x, pa["gn"][0], pa["gn"][1], rows_per_sample=hw, eps=1e-6, silu=False, chan_sums=hs)
        qk = ops.linear(xn, pa["w_qk"], pa["b_qk"])                      # [n*hw, 2c]
        q = qk[:, :c].view(n, hw, c)                                      # strided views of the fused projection
        k = qk[:, c:].view(n, hw, c)
        s = ops.bmm_nt(q, k)                                              # [n, hw, hw]
        ops.softmax_rows_(s, float(c) ** -0.5)
        # V^T for all frames in one GEMM: [c, n*hw] = W_v @ xn^T
        vt = ops.linear(pa["w_v"], xn, None)                              # A = W_v [c, c], B = xn [n*hw, c]
        o = ops.bmm_nt(s, vt.view(c, n, hw).permute(1, 0, 2))             # [n, hw, c] = P @ V (V^T read in place)
        so = pool.take(n, c, c, (hw, n, 1, 1), sdims=(1,))
        return ops.linear_frames(o.view(-1, c), pa["w_o"], pa["b_o"], hw=hw, residual=x, stats=so).view(n, hh, ww, c), so

    # ------------------------------------------------------------------ API
    @torch.no_grad()
    def decode_frames(self, z, scale=1.0):
        """z: [B, C, T, h, w] latent (any float dtype) -> [B, 3, T, 8h, 8w] in z.dtype; applies
        `scale * z`, post_quant_conv and the decoder with all B*T frames batched."""
        if not z.is_cuda:
            raise RuntimeError("AutoencoderKL(B200): input must be a CUDA tensor (no CPU fallback)")
        if self._packed is None:
            self.pack()
        P = self._packed
        b, c, t, hh, ww = z.shape
        z = z.contiguous()
        fr = torch.empty((b * t, hh, ww, c), device=z.device, dtype=BF16)
        ops._launch("bcthw_to_frames_mix", 0, lib().t2v_bcthw_to_frames_mix, z.data_ptr(), DTYPE_CODE[z.dtype],
                    fr.data_ptr(), b, c, t, hh, ww, float(scale), P["pq_w"].data_ptr(), P["pq_b"].data_ptr(), stream_ptr())
        w, bias, cout = P["conv_in"]
        h = ops.conv3x3_small_cin(fr, w, bias, cout)
        pool = _StatsPool(z.device, b * t * 512 * 2 * 48)
        h, hs = self._res(P["mid1"], h, None, pool)       # (conv_in is a direct 4-channel conv: no sums for the first norm)
        h, hs = self._attn(P["attn"], h, hs, pool)
        h, hs = self._res(P["mid2"], h, hs, pool)
        for i_level in reversed(range(len(P["up"]))):
            blocks, ups = P["up"][i_level]
            for pr in blocks:
                h, hs = self._res(pr, h, hs, pool)
            if ups is not None:
                hs = pool.take(h.shape[0], ups[0].shape[1], ups[0].shape[2], (h.shape[2], h.shape[1], h.shape[0], 1))
                h = ops.upconv3x3(h, ups[0], ups[1], stats=hs)
        n, hh2, ww2, ch = h.shape
        hn = ops.groupnorm(h.view(-1, ch), P["norm_out"][0], P["norm_out"][1], rows_per_sample=hh2 * ww2, eps=1e-6, silu=True,
                           chan_sums=hs)
        w, bias, cout = P["conv_out"]
        y = ops.conv3x3(hn.view(n, hh2, ww2, ch), w, bias, bias_div=n)
        return ops.frames_to_bcthw(y, b, cout, z.dtype)

    @torch.no_grad()
    def _encode_moments(self, x):
        """x: [B, 3, T, H, W] -> fp32 channels-last moments [B*T, H/f, W/f, 2*embed] (mean | logvar) of the posterior:
        Encoder (ae_modules.py:470-503) and quant_conv (autoencoder.py:105-106), all frames batched."""
        if not x.is_cuda:
            raise RuntimeError("AutoencoderKL(B200): input must be a CUDA tensor (no CPU fallback)")
        if self._packed_enc is None:
            self.pack_encoder()
        P = self._packed_enc
        b, c, t, hh, ww = x.shape
        x = x.contiguous()
        fr = torch.empty((b * t, hh, ww, 4), device=x.device, dtype=BF16)
        ops._launch("bcthw_to_frames_pad", 0, lib().t2v_bcthw_to_frames_pad, x.data_ptr(), DTYPE_CODE[x.dtype],
                    fr.data_ptr(), b, c, 4, t, hh, ww, 1.0, stream_ptr())
        w, bias, cout = P["conv_in"]
        h = ops.conv3x3_small_cin(fr, w, bias, cout)
        pool = _StatsPool(x.device, b * t * 512 * 2 * 48)
        hs = None
        for blocks, ds in P["down"]:
            for pr in blocks:
                h, hs = self._res(pr, h, hs, pool)
            if ds is not None:
                hs = pool.take(h.shape[0], ds[0].shape[0], ds[0].shape[1], (h.shape[2] // 2, 1, h.shape[1] // 2, h.shape[0]), sdims=(3,))
                h = ops.conv3x3_s2(h, ds[0], ds[1], pad="br", stats=hs)
        h, hs = self._res(P["mid1"], h, hs, pool)
        h, hs = self._attn(P["attn"], h, hs, pool)
        h, hs = self._res(P["mid2"], h, hs, pool)
        n, h2, w2, ch = h.shape
        hn = ops.groupnorm(h.view(-1, ch), P["norm_out"][0], P["norm_out"][1], rows_per_sample=h2 * w2, eps=1e-6, silu=True,
                           chan_sums=hs)
        w, bias, cout = P["conv_out"]
        return ops.conv3x3(hn.view(n, h2, w2, ch), w, bias, bias_div=n, out_f32=True)   # fp32 [n, h, w, 2*embed]

    @torch.no_grad()
    def encode_frames(self, x, noise=None, scale=1.0, sample=True):
        """x: [B, 3, T, H, W] video (any float dtype) -> scale * posterior sample (or mode) [B, zc, T, H/f, W/f] in
        x.dtype, all B*T frames batched (ddpm3d.py:558-584).  noise: fp32 [B*T, zc, h, w] (the reference draws it
        with torch.randn on the CPU, distributions.py:38-41 — done here the same way when omitted and sample=True)."""
        b, _, t = x.shape[:3]
        moments = self._encode_moments(x)
        n, h2, w2, c2 = moments.shape
        zc = c2 // 2
        if sample and noise is None:
            noise = torch.randn((n, zc, h2, w2))
        return ops.gaussian_sample(moments, noise if sample else None, b=b, t=t, zc=zc, scale=scale, dtype=x.dtype)

    def encode(self, x, **kwargs):
        """autoencoder.py:103-108: x [N, 3, H, W] -> the posterior; callers do `.sample()` / `.mode()`
        (train_t2v_turbo_v1_lora.py:958-965, ddpm3d.py:558-567)."""
        return DiagonalGaussianDistribution(self._encode_moments(x.unsqueeze(2)), x.dtype)

    def decode(self, z, **kwargs):
        """autoencoder.py:110-113: z [N, C, h, w] -> [N, 3, 8h, 8w]."""
        return self.decode_frames(z.unsqueeze(2), 1.0).squeeze(2)
