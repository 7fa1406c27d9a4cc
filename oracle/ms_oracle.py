"""ORACLE (test infrastructure) for the ModelScope path (BASELINE config 5) — PARITY UNPINNED.

`UNet3DConditionModel.forward` (model_scope/unet_3d_condition.py:329-504) wires diffusers blocks
(model_scope/unet_3d_blocks.py:268-875).  diffusers is pinned at 0.30.0 by the reference (cog.yaml:14) but is NOT installed
here and is absent from /root/reference, so its arithmetic cannot be executed: this file RESTATES, in plain fp32 torch and
under diffusers' own state-dict key names, the published algorithms of
  models/resnet.py            ResnetBlock2D (pre-norm, swish, time_embedding_norm="default", output_scale_factor 1),
                              TemporalConvLayer (4 x [GroupNorm32 -> SiLU -> (Dropout) -> Conv3d (3,1,1)], residual),
                              Downsample2D (conv 3x3 stride 2 pad 1), Upsample2D (nearest 2x -> conv 3x3)
  models/transformers/        Transformer2DModel (GroupNorm eps 1e-6 -> Linear proj_in -> BasicTransformerBlock -> proj_out, +x),
                              TransformerTemporalModel (same over "(b hw) f c" sequences, double self-attention)
  models/attention.py         BasicTransformerBlock (LayerNorm -> attn1 self, -> attn2, -> GEGLU feed-forward), Attention
  models/embeddings.py        Timesteps(320, flip_sin_to_cos=True, freq_shift=0), TimestepEmbedding(cond_proj_dim=256)
anchored on the reference's call sites (block wiring and argument values in unet_3d_blocks.py / unet_3d_condition.py).
The reference has no test or golden vector at this boundary.  `ms_unet_param_shapes` is an independent census of the
parameters (name -> shape) written from those constructors; the B200 key map is tested against it.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from oracle.unet_oracle import _SD, _gn, _linear, spatial_transformer, temporal_conv_block, temporal_transformer


def timesteps_proj(t, dim):
    """embeddings.get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin], exponent / half."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def resnet_block_2d(p: _SD, x, temb):
    h = F.silu(_gn(p.sub("norm1"), x, 1e-5))
    h = F.conv2d(h, p["conv1.weight"], p["conv1.bias"], padding=1)
    h = h + _linear(p.sub("time_emb_proj"), F.silu(temb))[:, :, None, None]
    h = F.silu(_gn(p.sub("norm2"), h, 1e-5))
    h = F.conv2d(h, p["conv2.weight"], p["conv2.bias"], padding=1)
    if p.has("conv_shortcut.weight"):
        x = F.conv2d(x, p["conv_shortcut.weight"], p["conv_shortcut.bias"])
    return x + h


def _frames5(h, b):
    bt, c, hh, ww = h.shape
    return h.reshape(b, bt // b, c, hh, ww).permute(0, 2, 1, 3, 4)


def _frames4(h5):
    b, c, t, hh, ww = h5.shape
    return h5.permute(0, 2, 1, 3, 4).reshape(b * t, c, hh, ww)


def unet3d_forward(sd, cfg, sample, timestep, encoder_hidden_states, timestep_cond=None):
    """sd: diffusers-keyed fp32 state dict; cfg: dict(block_out_channels, layers_per_block); sample [B, 4, F, h, w]."""
    p = _SD(sd)
    boc, lpb = tuple(cfg["block_out_channels"]), cfg["layers_per_block"]
    n_levels = len(boc)
    b, _, f, _, _ = sample.shape
    t_emb = timesteps_proj(timestep.reshape(-1).expand(b), boc[0])
    if timestep_cond is not None:
        t_emb = t_emb + F.linear(timestep_cond.float(), sd["time_embedding.cond_proj.weight"])
    emb = _linear(p.sub("time_embedding.linear_2"), F.silu(_linear(p.sub("time_embedding.linear_1"), t_emb)))
    emb = emb.repeat_interleave(f, dim=0)
    ctx = encoder_hidden_states.float().repeat_interleave(f, dim=0)
    h = sample.float().permute(0, 2, 1, 3, 4).reshape(b * f, sample.shape[1], sample.shape[3], sample.shape[4])
    h = F.conv2d(h, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)
    h = _frames4(temporal_transformer(p.sub("transformer_in"), _frames5(h, b), 8))

    def layer(q: _SD, j, h, attn, ch):
        h = resnet_block_2d(q.sub(f"resnets.{j}"), h, emb)
        h = _frames4(temporal_conv_block(q.sub(f"temp_convs.{j}"), _frames5(h, b)))
        if attn:
            h = spatial_transformer(q.sub(f"attentions.{j}"), h, ctx, ch // 64)
            h = _frames4(temporal_transformer(q.sub(f"temp_attentions.{j}"), _frames5(h, b), ch // 64))
        return h
    res = [h]
    for lvl in range(n_levels):
        q = p.sub(f"down_blocks.{lvl}")
        for j in range(lpb):
            h = layer(q, j, h, lvl < n_levels - 1, boc[lvl])
            res.append(h)
        if lvl < n_levels - 1:
            h = F.conv2d(h, q["downsamplers.0.conv.weight"], q["downsamplers.0.conv.bias"], stride=2, padding=1)
            res.append(h)
    q = p.sub("mid_block")
    h = resnet_block_2d(q.sub("resnets.0"), h, emb)
    h = _frames4(temporal_conv_block(q.sub("temp_convs.0"), _frames5(h, b)))
    h = spatial_transformer(q.sub("attentions.0"), h, ctx, boc[-1] // 64)
    h = _frames4(temporal_transformer(q.sub("temp_attentions.0"), _frames5(h, b), boc[-1] // 64))
    h = resnet_block_2d(q.sub("resnets.1"), h, emb)
    h = _frames4(temporal_conv_block(q.sub("temp_convs.1"), _frames5(h, b)))
    rev = boc[::-1]
    for u in range(n_levels):
        q = p.sub(f"up_blocks.{u}")
        for j in range(lpb + 1):
            h = torch.cat([h, res.pop()], dim=1)
            h = layer(q, j, h, u > 0, rev[u])
        if u < n_levels - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, q["upsamplers.0.conv.weight"], q["upsamplers.0.conv.bias"], padding=1)
    h = F.silu(_gn(p.sub("conv_norm_out"), h, 1e-5))
    h = F.conv2d(h, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)
    return _frames5(h, b)


def ms_unet_param_shapes(block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, in_channels=4, out_channels=4,
                         cross_attention_dim=1024, time_cond_proj_dim=256):
    """name -> shape of every parameter of UNet3DConditionModel, from unet_3d_condition.py:87-283 and the constructors of the
    diffusers blocks it instantiates (independent of the B200 key map)."""
    boc = tuple(block_out_channels)
    ted = boc[0] * 4
    s = {}

    def lin(n, i, o, bias=True):
        s[f"{n}.weight"] = (o, i)
        if bias:
            s[f"{n}.bias"] = (o,)

    def norm(n, c):
        s[f"{n}.weight"], s[f"{n}.bias"] = (c,), (c,)

    def conv(n, i, o, k):
        s[f"{n}.weight"], s[f"{n}.bias"] = (o, i) + k, (o,)

    def resnet(n, i, o):
        norm(f"{n}.norm1", i); conv(f"{n}.conv1", i, o, (3, 3)); lin(f"{n}.time_emb_proj", ted, o)
        norm(f"{n}.norm2", o); conv(f"{n}.conv2", o, o, (3, 3))
        if i != o:
            conv(f"{n}.conv_shortcut", i, o, (1, 1))

    def temp_conv(n, c):
        for k, idx in ((1, 2), (2, 3), (3, 3), (4, 3)):
            norm(f"{n}.conv{k}.0", c); conv(f"{n}.conv{k}.{idx}", c, c, (3, 1, 1))

    def block(n, dim, ctx):     # BasicTransformerBlock: attention_bias False, GEGLU
        for k in ("norm1", "norm2", "norm3"):
            norm(f"{n}.{k}", dim)
        for a, kv in (("attn1", dim), ("attn2", ctx)):
            lin(f"{n}.{a}.to_q", dim, dim, False); lin(f"{n}.{a}.to_k", kv, dim, False); lin(f"{n}.{a}.to_v", kv, dim, False)
            lin(f"{n}.{a}.to_out.0", dim, dim)
        lin(f"{n}.ff.net.0.proj", dim, dim * 8); lin(f"{n}.ff.net.2", dim * 4, dim)

    def transformer(n, c, inner, ctx):
        norm(f"{n}.norm", c); lin(f"{n}.proj_in", c, inner); block(f"{n}.transformer_blocks.0", inner, ctx); lin(f"{n}.proj_out", inner, c)
    conv("conv_in", in_channels, boc[0], (3, 3))
    lin("time_embedding.linear_1", boc[0], ted); lin("time_embedding.linear_2", ted, ted)
    lin("time_embedding.cond_proj", time_cond_proj_dim, boc[0], False)
    transformer("transformer_in", boc[0], 512, 512)                  # TransformerTemporalModel(8 x 64), double self-attention
    n_levels = len(boc)
    ch = boc[0]
    skips = [ch]
    for lvl in range(n_levels):
        for j in range(layers_per_block):
            resnet(f"down_blocks.{lvl}.resnets.{j}", ch, boc[lvl]); ch = boc[lvl]
            temp_conv(f"down_blocks.{lvl}.temp_convs.{j}", ch)
            if lvl < n_levels - 1:
                transformer(f"down_blocks.{lvl}.attentions.{j}", ch, ch, cross_attention_dim)
                transformer(f"down_blocks.{lvl}.temp_attentions.{j}", ch, ch, ch)
            skips.append(ch)
        if lvl < n_levels - 1:
            conv(f"down_blocks.{lvl}.downsamplers.0.conv", ch, ch, (3, 3)); skips.append(ch)
    for j in (0, 1):
        resnet(f"mid_block.resnets.{j}", ch, ch); temp_conv(f"mid_block.temp_convs.{j}", ch)
    transformer("mid_block.attentions.0", ch, ch, cross_attention_dim); transformer("mid_block.temp_attentions.0", ch, ch, ch)
    rev = boc[::-1]
    for u in range(n_levels):
        for j in range(layers_per_block + 1):
            resnet(f"up_blocks.{u}.resnets.{j}", ch + skips.pop(), rev[u]); ch = rev[u]
            temp_conv(f"up_blocks.{u}.temp_convs.{j}", ch)
            if u > 0:
                transformer(f"up_blocks.{u}.attentions.{j}", ch, ch, cross_attention_dim)
                transformer(f"up_blocks.{u}.temp_attentions.{j}", ch, ch, ch)
        if u < n_levels - 1:
            conv(f"up_blocks.{u}.upsamplers.0.conv", ch, ch, (3, 3))
    norm("conv_norm_out", boc[0]); conv("conv_out", boc[0], out_channels, (3, 3))
    return s
