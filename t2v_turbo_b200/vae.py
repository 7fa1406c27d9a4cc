"""Drop-in KL-VAE decode path on B200: `AutoencoderKL.decode` / `decode_first_stage_2DAE`.

Mirrors lvdm/models/autoencoder.py:110-113 and lvdm/modules/networks/ae_modules.py:506-641 (Decoder,
ResnetBlock :146-203, AttnBlock :29-73, Upsample :108-122): same constructor config (`ddconfig`,
`embed_dim`), same state-dict keys (`post_quant_conv.*`, `decoder.conv_in.*`, `decoder.mid.block_1.*`,
`decoder.up.{i}.block.{j}.*`, `decoder.up.{i}.upsample.conv.*`, ...).  The nn modules are parameter
containers; the arithmetic is libt2v_b200.so kernels over channels-last bf16 with ALL frames batched
(the reference decodes frame by frame in a Python loop, ddpm3d.py:671-677).
The encoder (training only) is out of scope this round.
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
    """Zeroed fp32 workspace for the per-frame channel sums the GEMM epilogues accumulate (one memset per decode)."""

    def __init__(self, device, nfloats):
        self.buf = torch.zeros((nfloats,), device=device, dtype=torch.float32)
        self.cur = 0

    def take(self, frames, channels, k_total, grid):
        if not ops.gn_fuse_producer(k_total, grid):
            return None
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


class AutoencoderKL(nn.Module):
    """Decode-side AutoencoderKL (reference ctor: ddconfig, lossconfig, embed_dim; lossconfig ignored)."""

    def __init__(self, ddconfig, embed_dim, lossconfig=None, **ignored):
        super().__init__()
        self.ddconfig, self.embed_dim = dict(ddconfig), embed_dim
        self.decoder = Decoder(**ddconfig)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        self._packed = None

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

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
        a = d.mid.attn_1
        c = a.in_channels
        wo = a.proj_out.weight.detach().float().reshape(c, c)
        # v bias folded through the softmax (rows sum to 1) into the output projection bias
        P["attn"] = dict(gn=(_f32(a.norm.weight), _f32(a.norm.bias)),
                         w_qk=torch.cat([_w2d(a.q.weight), _w2d(a.k.weight)], 0).contiguous(),
                         b_qk=torch.cat([_f32(a.q.bias), _f32(a.k.bias)], 0).contiguous(),
                         w_v=_w2d(a.v.weight), w_o=_w2d(a.proj_out.weight),
                         b_o=(_f32(a.proj_out.bias) + wo @ _f32(a.v.bias)).contiguous(), c=c)
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
        so = pool.take(n, c, c, (hw, n, 1, 1))
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

    def decode(self, z, **kwargs):
        """autoencoder.py:110-113: z [N, C, h, w] -> [N, 3, 8h, 8w]."""
        return self.decode_frames(z.unsqueeze(2), 1.0).squeeze(2)

    def encode(self, x, **kwargs):
        raise NotImplementedError("AutoencoderKL(B200): the encoder (training-time only) is not built in this round")
