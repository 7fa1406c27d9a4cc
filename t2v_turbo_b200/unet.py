"""Drop-in `UNetModel` for the VideoCrafter2 / T2V-Turbo latent-video UNet on B200.

Contract mirrored from the reference (lvdm/modules/networks/openaimodel3d.py:312-740,
lvdm/modules/attention.py:50-542): same constructor kwargs, same state-dict key names and shapes
(`input_blocks.N.M...`, `temopral_conv` [sic], `transformer_blocks.0.attn1.to_q.weight`, ...), same
`forward(x, timesteps, context=..., fps=..., timestep_cond=..., motion_cond=...)` signature and
`[B, C, T, H, W]` in / out.  The torch.nn modules below are PARAMETER CONTAINERS only: `forward`
never calls them.  The arithmetic is a fixed sequence of libt2v_b200.so kernels over ONE
channels-last bf16 layout `[B*T, H, W, C]` (see ops.py), with weights pre-packed once
(bf16, tap-major conv kernels, fused QKV, GEGLU row interleave, folded embedding biases).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops

BF16 = torch.bfloat16


# =============================================================================== containers
def _gn(ch, eps=1e-5):
    return nn.GroupNorm(32, ch, eps=eps)


class _Seq(nn.Sequential):
    """Sequential used only for its numbered children (reference: TimestepEmbedSequential)."""


class Downsample(nn.Module):
    def __init__(self, channels, out_channels=None):
        super().__init__()
        self.channels, self.out_channels = channels, out_channels or channels
        self.op = nn.Conv2d(channels, self.out_channels, 3, stride=2, padding=1)


class Upsample(nn.Module):
    def __init__(self, channels, out_channels=None):
        super().__init__()
        self.channels, self.out_channels = channels, out_channels or channels
        self.conv = nn.Conv2d(channels, self.out_channels, 3, padding=1)


class TemporalConvBlock(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv1 = nn.Sequential(_gn(ch), nn.SiLU(), nn.Conv3d(ch, ch, (3, 1, 1), padding=(1, 0, 0)))
        for i in (2, 3, 4):
            setattr(self, f"conv{i}", nn.Sequential(_gn(ch), nn.SiLU(), nn.Dropout(0.1),
                                                    nn.Conv3d(ch, ch, (3, 1, 1), padding=(1, 0, 0))))
        nn.init.zeros_(self.conv4[-1].weight)
        nn.init.zeros_(self.conv4[-1].bias)


class ResBlock(nn.Module):
    def __init__(self, channels, emb_channels, out_channels=None, use_temporal_conv=False):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_temporal_conv = use_temporal_conv
        oc = self.out_channels
        self.in_layers = nn.Sequential(_gn(channels), nn.SiLU(), nn.Conv2d(channels, oc, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, oc))
        out_conv = nn.Conv2d(oc, oc, 3, padding=1)
        nn.init.zeros_(out_conv.weight)
        nn.init.zeros_(out_conv.bias)
        self.out_layers = nn.Sequential(_gn(oc), nn.SiLU(), nn.Dropout(0.0), out_conv)
        self.skip_connection = nn.Identity() if oc == channels else nn.Conv2d(channels, oc, 1)
        if use_temporal_conv:
            self.temopral_conv = TemporalConvBlock(oc)  # (sic) reference spelling, openaimodel3d.py:196


class CrossAttention(nn.Module):
    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, record_attn_probs=False):
        super().__init__()
        self.record_attn_probs = record_attn_probs   # attention.py:99-100: keep softmax(q k^T) of the last forward
        self.attention_probs = None
        inner = heads * dim_head
        context_dim = context_dim or query_dim
        self.heads, self.dim_head, self.scale = heads, dim_head, dim_head ** -0.5
        self.is_self = context_dim == query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(0.0))


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.Sequential(GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, n_heads, d_head, context_dim=None, record_attn_probs=False):
        super().__init__()
        self.attn1 = CrossAttention(dim, None, n_heads, d_head, record_attn_probs=record_attn_probs)   # attention.py:262-269
        self.ff = FeedForward(dim)
        self.attn2 = CrossAttention(dim, context_dim, n_heads, d_head)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)


class SpatialTransformer(nn.Module):
    def __init__(self, in_channels, n_heads, d_head, depth=1, context_dim=None, use_linear=False):
        super().__init__()
        assert depth == 1, "transformer_depth > 1 is not used by VC2 / T2V-Turbo"
        inner = n_heads * d_head
        self.in_channels, self.use_linear = in_channels, use_linear
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6)
        self.proj_in = nn.Linear(in_channels, inner) if use_linear else nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(inner, n_heads, d_head, context_dim)])
        self.proj_out = nn.Linear(inner, in_channels) if use_linear else nn.Conv2d(inner, in_channels, 1)
        nn.init.zeros_(self.proj_out.weight)
        nn.init.zeros_(self.proj_out.bias)


class TemporalTransformer(nn.Module):
    def __init__(self, in_channels, n_heads, d_head, depth=1, use_linear=False, only_self_att=True, record_attn_probs=False):
        super().__init__()
        assert depth == 1 and only_self_att, "VC2 config: temporal_selfatt_only, depth 1"
        inner = n_heads * d_head
        self.in_channels, self.use_linear = in_channels, use_linear
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6)
        self.proj_in = nn.Linear(in_channels, inner) if use_linear else nn.Conv1d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(inner, n_heads, d_head, None, record_attn_probs=record_attn_probs)])
        self.proj_out = nn.Linear(inner, in_channels) if use_linear else nn.Conv1d(inner, in_channels, 1)
        nn.init.zeros_(self.proj_out.weight)
        nn.init.zeros_(self.proj_out.bias)


# =============================================================================== packed weights
def _w2d(w):
    """Linear / 1x1-conv weight -> bf16 [N, K]."""
    return w.detach().reshape(w.shape[0], -1).to(BF16).contiguous()


def _f32(t):
    return t.detach().float().contiguous()


import os as _os
_DBG_TAGS = set(_os.environ["T2V_GN_FUSE_TAGS"].split(",")) if "T2V_GN_FUSE_TAGS" in _os.environ else None   # debugging only


class _PackedAttn:
    """to_q (| to_k | to_v) with the preceding LayerNorm folded in (ops.fold_layernorm): the projection GEMM reads the
    un-normalised residual stream and its epilogue applies the per-row statistics."""

    def __init__(self, attn: CrossAttention, fuse_qkv: bool, norm: nn.LayerNorm):
        self.heads, self.scale = attn.heads, attn.scale
        self.inner = attn.heads * attn.dim_head
        ws = [attn.to_q.weight] + ([attn.to_k.weight, attn.to_v.weight] if fuse_qkv else [])
        w = torch.cat([x.detach().float() for x in ws], 0)
        self.w_qkv, self.b_qkv, self.cs_qkv = ops.fold_layernorm(w, None, norm.weight.detach(), norm.bias.detach())
        self.fused = fuse_qkv
        self.w_o, self.b_o = _w2d(attn.to_out[0].weight), _f32(attn.to_out[0].bias)


class _PackedBlock:
    def __init__(self, blk: BasicTransformerBlock, cross: bool):
        self.a1 = _PackedAttn(blk.attn1, True, blk.norm1)
        self.a2 = _PackedAttn(blk.attn2, not cross, blk.norm2)
        self.ln_eps = [n.eps for n in (blk.norm1, blk.norm2, blk.norm3)]
        proj = blk.ff.net[0].proj
        wf, bf, cs = ops.fold_layernorm(proj.weight, proj.bias, blk.norm3.weight.detach(), blk.norm3.bias.detach())
        self.w_ff1, self.b_ff1 = ops.pack_geglu(wf, bf)
        _, self.cs_ff1 = ops.pack_geglu(wf, cs)
        self.w_ff2, self.b_ff2 = _w2d(blk.ff.net[2].weight), _f32(blk.ff.net[2].bias)


class _PackedTransformer:
    def __init__(self, m, temporal: bool):
        self.temporal = temporal
        self.gn = (_f32(m.norm.weight), _f32(m.norm.bias), m.norm.eps)
        self.w_in, self.b_in = _w2d(m.proj_in.weight), _f32(m.proj_in.bias)
        self.w_out, self.b_out = _w2d(m.proj_out.weight), _f32(m.proj_out.bias)
        self.blk = _PackedBlock(m.transformer_blocks[0], cross=not temporal)
        self.kv_slice = None  # (offset, inner) into the batched context K/V projection
        a1 = m.transformer_blocks[0].attn1
        self.record_attn = a1 if (temporal and getattr(a1, "record_attn_probs", False)) else None


class _PackedRes:
    def __init__(self, rb: ResBlock):
        g = rb.in_layers[0]
        self.gn1 = (_f32(g.weight), _f32(g.bias), g.eps)
        self.w1 = ops.pack_conv_weight(rb.in_layers[2].weight.detach())
        g = rb.out_layers[0]
        self.gn2 = (_f32(g.weight), _f32(g.bias), g.eps)
        self.w2, self.b2 = ops.pack_conv_weight(rb.out_layers[3].weight.detach()), _f32(rb.out_layers[3].bias).view(1, -1)
        if isinstance(rb.skip_connection, nn.Identity):
            self.w_skip = None
        else:
            self.w_skip, self.b_skip = _w2d(rb.skip_connection.weight), _f32(rb.skip_connection.bias)
        self.cout = rb.out_channels
        self.emb_slice = None
        self.tconv = None
        if rb.use_temporal_conv:
            self.tconv = []
            for i in (1, 2, 3, 4):
                seq = getattr(rb.temopral_conv, f"conv{i}")
                self.tconv.append(((_f32(seq[0].weight), _f32(seq[0].bias), seq[0].eps),
                                   ops.pack_conv_weight(seq[-1].weight.detach()), _f32(seq[-1].bias)))


# =============================================================================== the model
class UNetModel(nn.Module):
    """B200-native VideoCrafter2 UNet (see module docstring). Unused reference kwargs are accepted."""

    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0.0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, context_dim=None,
                 use_scale_shift_norm=False, resblock_updown=False, num_heads=-1, num_head_channels=-1,
                 transformer_depth=1, use_linear=False, use_checkpoint=False, temporal_conv=False,
                 tempspatial_aware=False, temporal_attention=True, temporal_selfatt_only=True,
                 use_relative_position=True, use_causal_attention=False, temporal_length=None, use_fp16=False,
                 addition_attention=False, use_image_attention=False, temporal_transformer_depth=1,
                 fps_cond=False, time_cond_proj_dim=None, motion_cond_proj_dim=None, record_attn_probs=False):
        super().__init__()
        unsupported = dict(use_scale_shift_norm=use_scale_shift_norm, resblock_updown=resblock_updown,
                           tempspatial_aware=tempspatial_aware, use_relative_position=use_relative_position,
                           use_causal_attention=use_causal_attention, use_image_attention=use_image_attention)
        bad = [k for k, v in unsupported.items() if v]
        if bad or dims != 2 or not conv_resample or num_head_channels != 64 or not temporal_attention:
            raise NotImplementedError(f"UNetModel(B200): options outside the VC2/T2V-Turbo config: {bad}")
        self.in_channels, self.model_channels, self.out_channels = in_channels, model_channels, out_channels
        self.num_res_blocks, self.attention_resolutions = num_res_blocks, attention_resolutions
        self.channel_mult, self.temporal_attention = channel_mult, temporal_attention
        self.use_checkpoint = use_checkpoint
        self.dtype = torch.float16 if use_fp16 else torch.float32  # reference attribute; callers overwrite it
        self.addition_attention, self.fps_cond = addition_attention, fps_cond
        self.time_cond_proj_dim, self.motion_cond_proj_dim = time_cond_proj_dim, motion_cond_proj_dim
        self.context_dim = context_dim
        mc = model_channels
        ted = mc * 4

        def mlp():
            return nn.Sequential(nn.Linear(mc, ted), nn.SiLU(), nn.Linear(ted, ted))
        self.time_embed = mlp()
        if fps_cond:
            self.fps_embedding = mlp()
        self.time_cond_proj = nn.Linear(time_cond_proj_dim, mc, bias=False) if time_cond_proj_dim is not None else None
        if motion_cond_proj_dim is not None:
            self.motion_cond_proj = nn.Linear(motion_cond_proj_dim, mc, bias=False)
            self.combine_proj = nn.Linear(mc * 2, mc, bias=False)
        else:
            self.motion_cond_proj = self.combine_proj = None

        def res(cin, cout):
            return ResBlock(cin, ted, out_channels=cout, use_temporal_conv=temporal_conv)

        def attn_layers(ch, record=False):
            heads = ch // num_head_channels
            return [SpatialTransformer(ch, heads, num_head_channels, transformer_depth, context_dim, use_linear),
                    TemporalTransformer(ch, heads, num_head_channels, temporal_transformer_depth, use_linear,
                                        temporal_selfatt_only, record_attn_probs=record)]

        self.input_blocks = nn.ModuleList([_Seq(nn.Conv2d(in_channels, mc, 3, padding=1))])
        if addition_attention:
            self.init_attn = _Seq(TemporalTransformer(mc, 8, num_head_channels, transformer_depth, False,
                                                      temporal_selfatt_only))
        chans = [mc]
        ch, ds = mc, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [res(ch, mult * mc)]
                ch = mult * mc
                if ds in attention_resolutions:
                    layers += attn_layers(ch)
                self.input_blocks.append(_Seq(*layers))
                chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(_Seq(Downsample(ch, ch)))
                chans.append(ch)
                ds *= 2
        self.middle_block = _Seq(res(ch, ch), *attn_layers(ch), res(ch, ch))
        self.output_blocks = nn.ModuleList()
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = chans.pop()
                layers = [res(ch + ich, mult * mc)]
                ch = mult * mc
                if ds in attention_resolutions:
                    # only the decoder's temporal self-attentions record (openaimodel3d.py:630-646); the motion-prior code
                    # reads output_blocks.{3..11}.2.transformer_blocks.0.attn1.attention_probs (motion_prior_sample.py:40-56)
                    layers += attn_layers(ch, record_attn_probs)
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch, ch))
                    ds //= 2
                self.output_blocks.append(_Seq(*layers))
        out_conv = nn.Conv2d(mc, out_channels, 3, padding=1)
        nn.init.zeros_(out_conv.weight)
        nn.init.zeros_(out_conv.bias)
        self.out = nn.Sequential(_gn(ch), nn.SiLU(), out_conv)
        self._packed = None
        self._ctx_cache = None
        self.weight_generation = 0   # bumped whenever the packed weights are dropped: CUDA graphs key on it
        self._ln_state = None
        self._ln_states = {}     # per input geometry: captured CUDA graphs keep pointing at their workspace

    # ------------------------------------------------------------------ packing
    def invalidate_packed(self):
        self._packed = None
        self._ctx_cache = None
        self.weight_generation = getattr(self, "weight_generation", 0) + 1

    def _apply(self, fn, *a, **k):  # .to()/.cuda()/.half() change the parameters -> repack lazily
        self.invalidate_packed()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self.invalidate_packed()
        return super().load_state_dict(*a, **k)

    @torch.no_grad()
    def pack(self):
        """Pre-pack all weights for the kernels (once per weight update)."""
        dev = self.time_embed[0].weight.device
        if dev.type != "cuda":
            raise RuntimeError("UNetModel(B200) runs on a CUDA device only (no CPU fallback): call .cuda() first")
        P = {}
        half = self.model_channels // 2
        P["freqs"] = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half).to(dev)
        for name in ("time_embed", "fps_embedding"):
            if hasattr(self, name):
                seq = getattr(self, name)
                P[name] = (_w2d(seq[0].weight), _f32(seq[0].bias), _w2d(seq[2].weight), _f32(seq[2].bias))
        for name in ("time_cond_proj", "motion_cond_proj", "combine_proj"):
            m = getattr(self, name)
            P[name] = _w2d(m.weight) if m is not None else None
        c0 = self.input_blocks[0][0]
        P["conv_in"] = (ops.pack_conv_weight(c0.weight.detach()), _f32(c0.bias))
        res_list, st_list = [], []

        def pack_seq(seq):
            items = []
            for layer in seq:
                if isinstance(layer, ResBlock):
                    pr = _PackedRes(layer)
                    res_list.append((layer, pr))
                    items.append(("res", pr))
                elif isinstance(layer, SpatialTransformer):
                    pt = _PackedTransformer(layer, False)
                    st_list.append((layer, pt))
                    items.append(("st", pt))
                elif isinstance(layer, TemporalTransformer):
                    items.append(("tt", _PackedTransformer(layer, True)))
                elif isinstance(layer, Downsample):
                    items.append(("down", (ops.pack_conv_weight(layer.op.weight.detach()), _f32(layer.op.bias))))
                elif isinstance(layer, Upsample):
                    items.append(("up", (ops.pack_upconv_weight(layer.conv.weight), _f32(layer.conv.bias))))
                elif isinstance(layer, nn.Conv2d):
                    items.append(("conv_in", None))
                else:
                    raise TypeError(type(layer))
            return items
        P["input"] = [pack_seq(s) for s in self.input_blocks]
        P["init_attn"] = pack_seq(self.init_attn) if self.addition_attention else None
        P["middle"] = pack_seq(self.middle_block)
        P["output"] = [pack_seq(s) for s in self.output_blocks]
        # all ResBlock embedding projections as ONE small linear; conv-in bias folded in
        ws, bs, off = [], [], 0
        for rb, pr in res_list:
            ws.append(_w2d(rb.emb_layers[1].weight))
            bs.append(_f32(rb.emb_layers[1].bias) + _f32(rb.in_layers[2].bias))
            pr.emb_slice = (off, pr.cout)
            off += pr.cout
        P["emb_w"], P["emb_b"] = torch.cat(ws, 0).contiguous(), torch.cat(bs, 0).contiguous()
        # all cross-attention K/V projections of the text context as ONE GEMM
        ws, off = [], 0
        for st, pt in st_list:
            a2 = st.transformer_blocks[0].attn2
            ws.append(torch.cat([_w2d(a2.to_k.weight), _w2d(a2.to_v.weight)], 0))
            pt.kv_slice = (off, pt.blk.a2.inner)
            off += 2 * pt.blk.a2.inner
        P["ctx_w"] = torch.cat(ws, 0).contiguous() if ws else None
        g = self.out[0]
        P["out"] = ((_f32(g.weight), _f32(g.bias), g.eps), ops.pack_conv_weight(self.out[2].weight.detach()),
                    _f32(self.out[2].bias).view(1, -1))
        self._packed = P
        return self

    # ------------------------------------------------------------------ forward pieces
    def _embedding(self, P, timesteps, fps, timestep_cond, motion_cond, bsz, dev):
        rnd = self.dtype != torch.float32  # the reference casts the sinusoid to self.dtype
        t = timesteps.to(device=dev, dtype=torch.float32).reshape(-1)
        if t.numel() == 1 and bsz > 1:
            t = t.expand(bsz).contiguous()
        t_emb = ops.sinusoidal_embedding(t, P["freqs"], round_bf16=rnd)
        add = None
        if timestep_cond is not None:
            tc = timestep_cond.to(device=dev, dtype=torch.float32).contiguous()
            cond = ops.small_linear(tc, P["time_cond_proj"], None, round_bf16=rnd)
            if motion_cond is not None:
                mcnd = ops.small_linear(motion_cond.to(device=dev, dtype=torch.float32).contiguous(), P["motion_cond_proj"], None, round_bf16=rnd)
                cond = ops.small_linear(torch.cat([cond, mcnd], 1), P["combine_proj"], None, round_bf16=rnd)
            add = cond
        w0, b0, w2, b2 = P["time_embed"]
        if add is not None:
            t_emb = t_emb + add  # tiny [B, 320] fp32 add (torch elementwise on 320 values: plumbing)
            if rnd:
                t_emb = t_emb.to(BF16).float()
        h = ops.small_linear(t_emb, w0, b0, silu_out=True, round_bf16=rnd)
        emb = ops.small_linear(h, w2, b2, round_bf16=rnd)
        if self.fps_cond:
            if isinstance(fps, int) or (torch.is_tensor(fps) and fps.numel() == 1 and bsz > 1):
                fps_t = torch.full((bsz,), float(fps), device=dev, dtype=torch.float32)
            else:
                fps_t = fps.to(device=dev, dtype=torch.float32).reshape(-1)
            f_emb = ops.sinusoidal_embedding(fps_t, P["freqs"], round_bf16=rnd)
            w0, b0, w2, b2 = P["fps_embedding"]
            h = ops.small_linear(f_emb, w0, b0, silu_out=True, round_bf16=rnd)
            emb = ops.small_linear(h, w2, b2, add=emb, round_bf16=rnd)
        # per-ResBlock rows: Linear(SiLU(emb)) + emb bias + conv-in bias, all blocks in one launch
        return ops.small_linear(emb, P["emb_w"], P["emb_b"], silu_in=True, round_bf16=False)

    def _context_kv(self, P, context, dev):
        key = (context.data_ptr(), context._version, tuple(context.shape), context.dtype)
        capturing = torch.cuda.is_current_stream_capturing()   # a graph must contain the projection itself
        if not capturing and self._ctx_cache is not None and self._ctx_cache[0] == key:
            return self._ctx_cache[1]
        ctx = context.to(device=dev, dtype=BF16).reshape(-1, context.shape[-1]).contiguous()
        kv = ops.linear(ctx, P["ctx_w"], None).view(context.shape[0], context.shape[1], -1)
        if not capturing:
            self._ctx_cache = (key, kv, context)
        return kv

    # ---- statistics workspace.  LayerNorm row sums (fp32 [rows, 2]) and GroupNorm channel sums (fp32 [frames, C, 2]) are
    # accumulated by the epilogues of the GEMMs that produce the tensors; every instance gets its own slice of one
    # buffer that a single memset zeroes at the start of a forward.  The first forward of a geometry sizes the buffer
    # (individually zeroed tensors).
    def _ln_begin(self, key):
        st = self._ln_states.get(key)
        if st is None:
            st = self._ln_states[key] = dict(key=key, buf=None, need=0, cursor=0)
            self._ln_state = st
        elif st["buf"] is None and st["need"] > 0:
            self._ln_state = st
            st["buf"] = torch.empty((st["need"],), device=key[-1], dtype=torch.float32)
        self._ln_state = st
        if st["buf"] is not None:
            st["buf"].zero_()
            st["cursor"] = 0
        else:
            st["need"] = 0

    def _ws_alloc(self, nfloats, device):
        nfloats = (nfloats + 3) // 4 * 4   # 16-byte aligned slices (vector reductions)
        st = self._ln_state
        if st["buf"] is None:
            st["need"] += nfloats
            return torch.zeros((nfloats,), device=device, dtype=torch.float32)
        off = st["cursor"]
        st["cursor"] = off + nfloats
        return st["buf"][off:off + nfloats]

    def _ln_slice(self, rows, device):
        return self._ws_alloc(rows * 2, device)[:rows * 2].view(rows, 2)

    def _gn_slice(self, frames, channels, device, k_total, tag="", grid=None, fixed=(None, None, None, None), sdims=()):
        """Channel-sum slice for a producer GEMM with reduction length k_total, or None when the fusion does not pay
        (ops.gn_fuse_producer): the consuming GroupNorm then runs its own statistics pass."""
        if not ops.gn_fuse_producer(k_total, grid, fixed, sdims) or (_DBG_TAGS is not None and tag not in _DBG_TAGS):
            return None
        return self._ws_alloc(frames * channels * 2, device)[:frames * channels * 2].view(frames, channels, 2)

    def _block(self, blk: _PackedBlock, x, acc1, geom, temporal, ctx_kv, kv_slice, record=None):
        b, t, hh, ww = geom
        hw = hh * ww
        a1 = blk.a1
        inner = a1.inner
        # The three LayerNorms never run as kernels: the GEMM that produces each LayerNorm input accumulates the
        # per-row sum / sum of squares in its epilogue (row_accum), and the GEMM that consumes the normalised
        # activation applies them (ln=..., weights pre-multiplied by gamma: ops.fold_layernorm).
        c = x.shape[1]
        acc2, acc3 = self._ln_slice(x.shape[0], x.device), self._ln_slice(x.shape[0], x.device)
        qkv = ops.linear(x, a1.w_qkv, a1.b_qkv, ln=(acc1, a1.cs_qkv, (c, blk.ln_eps[0])))
        q, k, v = qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:]
        if temporal:
            probs = None
            if record is not None:   # "(b h) i j" softmax of the temporal self-attention, in the model dtype (attention.py:124-126)
                pd = self.dtype if self.dtype in (torch.float16, BF16) else torch.float32
                probs = torch.empty((b * hw * a1.heads, t, t), device=x.device, dtype=pd)
            att = ops.attention_temporal(q, k, v, b=b, t=t, hw=hw, heads=a1.heads, scale=a1.scale, probs=probs)
            if record is not None:
                record.attention_probs = probs
        else:
            att = ops.attention(q.view(b * t, hw, inner), k.view(b * t, hw, inner), v.view(b * t, hw, inner),
                                heads=a1.heads, scale=a1.scale).view(-1, inner)
        x = ops.linear(att, a1.w_o, a1.b_o, residual=x, row_accum=acc2)
        a2 = blk.a2
        ln2 = (acc2, a2.cs_qkv, (c, blk.ln_eps[1]))
        if temporal:
            qkv = ops.linear(x, a2.w_qkv, a2.b_qkv, ln=ln2)
            q, k, v = qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:]
            att = ops.attention_temporal(q, k, v, b=b, t=t, hw=hw, heads=a2.heads, scale=a2.scale)
        else:
            q = ops.linear(x, a2.w_qkv, a2.b_qkv, ln=ln2)
            off, kin = kv_slice
            kc, vc = ctx_kv[:, :, off:off + kin], ctx_kv[:, :, off + kin:off + 2 * kin]
            att = ops.attention(q.view(b * t, hw, inner), kc, vc, heads=a2.heads, scale=a2.scale,
                                kv_batch_div=t).view(-1, inner)
        x = ops.linear(att, a2.w_o, a2.b_o, residual=x, row_accum=acc3)
        g = ops.linear(x, blk.w_ff1, blk.b_ff1, geglu=True, ln=(acc3, blk.cs_ff1, (c, blk.ln_eps[2])))
        return ops.linear(g, blk.w_ff2, blk.b_ff2, residual=x)

    def _transformer(self, pt: _PackedTransformer, h, h_stats, geom, ctx_kv):
        """h_stats: per-frame channel sums of h from its producer (or None: GroupNorm runs its own statistics pass).
        Returns (out, per-frame channel sums of out)."""
        b, t, hh, ww = geom
        c = h.shape[-1]
        x_in = h.view(-1, c)
        rps = hh * ww * (t if pt.temporal else 1)
        xn = ops.groupnorm(x_in, pt.gn[0], pt.gn[1], rows_per_sample=rps, eps=pt.gn[2], silu=False,
                           chan_sums=h_stats if (not pt.temporal or ops.gn_fuse_temporal()) else None,
                           chan_group=t if pt.temporal else 1)
        acc1 = self._ln_slice(xn.shape[0], xn.device)
        x = ops.linear(xn, pt.w_in, pt.b_in, row_accum=acc1)
        x = self._block(pt.blk, x, acc1, geom, pt.temporal, ctx_kv, pt.kv_slice, record=pt.record_attn)
        so = self._gn_slice(b * t, c, h.device, c, "tr", (hh * ww, b * t, 1, 1), sdims=(1,))
        out = ops.linear_frames(x, pt.w_out, pt.b_out, hw=hh * ww, residual=x_in, stats=so)
        return out.view(b * t, hh, ww, c), so

    def _resblock(self, pr: _PackedRes, x, x_stats, emb_rows, geom):
        """x: tensor or (h, skip) pair; x_stats: matching per-frame channel sums (or None).  Every conv of the block
        accumulates the GroupNorm statistics of its output, so only the apply pass of each GroupNorm runs."""
        b, t, hh, ww = geom
        hw = hh * ww
        nf = b * t
        x0, x1 = x if isinstance(x, tuple) else (x, None)
        dev = x0.device
        cin = x0.shape[-1] + (x1.shape[-1] if x1 is not None else 0)
        hn = ops.groupnorm(x, pr.gn1[0], pr.gn1[1], rows_per_sample=hw, eps=pr.gn1[2], silu=True, chan_sums=x_stats)
        off, cout = pr.emb_slice
        rowbias = emb_rows[:, off:off + cout].contiguous()
        s1 = self._gn_slice(nf, cout, dev, 9 * cin, "res1", (ww, hh, nf, 1), sdims=(2,))
        h = ops.conv3x3(hn.view(nf, hh, ww, cin), pr.w1, rowbias, bias_div=t, stats=s1)
        hn2 = ops.groupnorm(h.view(-1, cout), pr.gn2[0], pr.gn2[1], rows_per_sample=hw, eps=pr.gn2[2], silu=True, chan_sums=s1)
        if pr.w_skip is None:
            res = x0
        else:
            xa = x0.view(-1, x0.shape[-1])
            xb = x1.view(-1, x1.shape[-1]) if x1 is not None else None
            res = ops.linear((xa, xb) if xb is not None else xa, pr.w_skip, pr.b_skip).view(nf, hh, ww, cout)
        s2 = self._gn_slice(nf, cout, dev, 9 * cout, "res2", (ww, hh, nf, 1), sdims=(2,))
        h = ops.conv3x3(hn2.view(nf, hh, ww, cout), pr.w2, pr.b2, bias_div=nf, residual=res, stats=s2)
        if pr.tconv is not None:
            ident = h.view(b, t, hw, cout)
            y, ys = ident, s2
            for i, (gn, w, bias) in enumerate(pr.tconv):
                yn = ops.groupnorm(y.view(-1, cout), gn[0], gn[1], rows_per_sample=t * hw, eps=gn[2], silu=True,
                                   chan_sums=ys if ops.gn_fuse_temporal() else None, chan_group=t)
                ys = self._gn_slice(nf, cout, dev, 3 * cout, "tconv", (hw, t, b, 1), sdims=(1, 2))
                y = ops.tconv3(yn.view(b, t, hw, cout), w, bias, residual=ident if i == 3 else None, stats=ys)
            h, s2 = y.view(nf, hh, ww, cout), ys
        return h, s2

    def _run_seq(self, items, h, h_stats, emb_rows, geom, ctx_kv, P):
        """h / h_stats: activation (or (h, skip) pair) and the per-frame channel sums its producer accumulated."""
        b, t, hh, ww = geom
        for kind, pk in items:
            if kind == "res":
                h, h_stats = self._resblock(pk, h, h_stats, emb_rows, geom)
            elif kind in ("st", "tt"):
                h, h_stats = self._transformer(pk, h, h_stats, geom, ctx_kv)
            elif kind == "down":
                h_stats = self._gn_slice(b * t, pk[0].shape[0], h.device, pk[0].shape[1], "down", (ww // 2, 1, hh // 2, b * t), (None, 1, None, None), sdims=(3,))
                h = ops.conv3x3_s2(h, pk[0], pk[1], stats=h_stats)
                hh, ww = hh // 2, ww // 2
                geom = (b, t, hh, ww)
            elif kind == "up":
                h_stats = self._gn_slice(b * t, pk[0].shape[1], h.device, pk[0].shape[2], "up", (ww, hh, b * t, 1), sdims=(2,))
                h = ops.upconv3x3(h, pk[0], pk[1], stats=h_stats)
                hh, ww = hh * 2, ww * 2
                geom = (b, t, hh, ww)
            elif kind == "conv_in":
                w, bias = P["conv_in"]
                h = ops.conv3x3_small_cin(h, w, bias, self.model_channels)
                h_stats = None   # 4-channel direct conv: the first GroupNorm computes its own statistics
        return h, h_stats, geom

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, x, timesteps, context=None, features_adapter=None, fps=16, timestep_cond=None,
                motion_cond=None, **kwargs):
        if features_adapter is not None:
            raise NotImplementedError("features_adapter is not used by T2V-Turbo")
        if not x.is_cuda:
            raise RuntimeError("UNetModel(B200): input must be a CUDA tensor (no CPU fallback)")
        if self._packed is None:
            self.pack()
        P = self._packed
        dev = x.device
        b, _, t, hh, ww = x.shape
        self._ln_begin((b, t, hh, ww, dev))
        emb_rows = self._embedding(P, timesteps, fps, timestep_cond, motion_cond, b, dev)
        ctx_kv = self._context_kv(P, context, dev) if P["ctx_w"] is not None else None
        h = ops.bcthw_to_frames(x, 1.0)
        geom = (b, t, hh, ww)
        hs = []
        st = None
        for i, items in enumerate(P["input"]):
            h, st, geom = self._run_seq(items, h, st, emb_rows, geom, ctx_kv, P)
            if i == 0 and P["init_attn"] is not None:
                h, st, geom = self._run_seq(P["init_attn"], h, st, emb_rows, geom, ctx_kv, P)
            hs.append((h, st))
        h, st, geom = self._run_seq(P["middle"], h, st, emb_rows, geom, ctx_kv, P)
        for items in P["output"]:
            skip, skip_st = hs.pop()
            pair_st = (st, skip_st) if (st is not None and skip_st is not None) else None
            h, st, geom = self._run_seq([items[0]], (h, skip), pair_st, emb_rows, geom, ctx_kv, P)
            h, st, geom = self._run_seq(items[1:], h, st, emb_rows, geom, ctx_kv, P)
        gn, w, bias = P["out"]
        c = h.shape[-1]
        hn = ops.groupnorm(h.view(-1, c), gn[0], gn[1], rows_per_sample=geom[2] * geom[3], eps=gn[2], silu=True, chan_sums=st)
        y = ops.conv3x3(hn.view(b * t, geom[2], geom[3], c), w, bias, bias_div=b * t)
        return ops.frames_to_bcthw(y, b, self.out_channels, x.dtype)
