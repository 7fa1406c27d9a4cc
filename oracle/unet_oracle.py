"""ORACLE (test infrastructure, never shipped): plain-torch restatement of the reference
VideoCrafter2 `UNetModel.forward`, in the reference's own NCHW layout and op order, driven by a
state dict with the reference's key names.  Computes in the dtype of the state dict (fp32 for
parity tests; bf16 to measure the reference's own bf16 deviation).

Pinned against the UNMODIFIED reference modules executed in the authoring container:
oracle/make_goldens.py runs `lvdm.modules.networks.openaimodel3d.UNetModel` on seeded weights /
inputs and stores the outputs under tests/golden/; tests/test_oracle.py checks this file against
those fixtures.  (The reference has no tests or golden vectors of its own — SURVEY.md §4.)

Each function cites the reference lines it restates.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg may import this module.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def timestep_embedding(timesteps, dim, max_period=10000):
    """lvdm/models/utils_diffusion.py:8-32 (repeat_only=False)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def guidance_scale_embedding(w, embedding_dim=256, dtype=torch.float32):
    """pipeline/t2v_turbo_vc2_pipeline.py:99-120 (get_w_embedding)."""
    w = w * 1000.0
    half = embedding_dim // 2
    emb = torch.log(torch.tensor(10000.0)) / (half - 1)
    emb = torch.exp(torch.arange(half, dtype=dtype) * -emb)
    emb = w.to(dtype)[:, None] * emb[None, :]
    return torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)


class _SD:
    """state-dict view with a key prefix."""

    def __init__(self, sd, prefix=""):
        self.sd, self.prefix = sd, prefix

    def sub(self, name):
        return _SD(self.sd, f"{self.prefix}{name}.")

    def __getitem__(self, name):
        return self.sd[self.prefix + name]

    def has(self, name):
        return (self.prefix + name) in self.sd


def _linear(p: _SD, x):
    return F.linear(x, p["weight"], p["bias"] if p.has("bias") else None)


def _gn(p: _SD, x, eps):
    # basics.py:78-89 GroupNormSpecific / nn.GroupNorm(32, C): stats in fp32, result in x.dtype
    return F.group_norm(x.float(), 32, p["weight"].float(), p["bias"].float(), eps).type(x.dtype)


def cross_attention(p: _SD, x, context, heads):
    """attention.py:102-164 CrossAttention.forward (no mask / relative position / image tokens)."""
    q = _linear(p.sub("to_q"), x)
    context = x if context is None else context
    k = _linear(p.sub("to_k"), context)
    v = _linear(p.sub("to_v"), context)
    b, n, inner = q.shape
    d = inner // heads

    def split(t):  # "b n (h d) -> (b h) n d"
        return t.reshape(t.shape[0], t.shape[1], heads, d).permute(0, 2, 1, 3).reshape(t.shape[0] * heads, t.shape[1], d)
    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum("bid,bjd->bij", q, k) * (d ** -0.5)
    sim = sim.softmax(dim=-1)
    out = torch.einsum("bij,bjd->bid", sim, v)
    out = out.reshape(b, heads, n, d).permute(0, 2, 1, 3).reshape(b, n, inner)
    return _linear(p.sub("to_out.0"), out)


def basic_transformer_block(p: _SD, x, context, heads):
    """attention.py:300-311: attn1 is self-attention; attn2 is cross (spatial) or self (temporal, context None)."""
    def ln(name, t):
        return F.layer_norm(t, (t.shape[-1],), p[f"{name}.weight"], p[f"{name}.bias"], 1e-5)
    x = cross_attention(p.sub("attn1"), ln("norm1", x), None, heads) + x
    x = cross_attention(p.sub("attn2"), ln("norm2", x), context, heads) + x
    h = _linear(p.sub("ff.net.0.proj"), ln("norm3", x))   # GEGLU attention.py:516-523
    a, gate = h.chunk(2, dim=-1)
    h = a * F.gelu(gate)
    return _linear(p.sub("ff.net.2"), h) + x


def spatial_transformer(p: _SD, x, context, heads):
    """attention.py:373-389 (use_linear=True): x [(b t), c, h, w], context [(b t), 77, 1024]."""
    b, c, h, w = x.shape
    x_in = x
    x = _gn(p.sub("norm"), x, 1e-6)
    x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
    x = _linear(p.sub("proj_in"), x)
    x = basic_transformer_block(p.sub("transformer_blocks.0"), x, context, heads)
    x = _linear(p.sub("proj_out"), x)
    x = x.reshape(b, h, w, c).permute(0, 3, 1, 2)
    return x + x_in


def temporal_transformer(p: _SD, x, heads):
    """attention.py:471-513 (only_self_att): x [b, c, t, h, w]; proj_in/out Linear or Conv1d(k=1)."""
    b, c, t, h, w = x.shape
    x_in = x
    x = _gn(p.sub("norm"), x, 1e-6)
    x = x.permute(0, 3, 4, 2, 1).reshape(b * h * w, t, c)          # "(b h w) t c"
    w_in = p["proj_in.weight"]
    x = F.linear(x, w_in.reshape(w_in.shape[0], -1), p["proj_in.bias"])   # Conv1d k=1 == Linear
    x = basic_transformer_block(p.sub("transformer_blocks.0"), x, None, heads)
    w_out = p["proj_out.weight"]
    x = F.linear(x, w_out.reshape(w_out.shape[0], -1), p["proj_out.bias"])
    x = x.reshape(b, h, w, t, c).permute(0, 4, 3, 1, 2)
    return x + x_in


def temporal_conv_block(p: _SD, x):
    """openaimodel3d.py:302-309: x + conv4(conv3(conv2(conv1(x)))), each GN32(eps 1e-5)->SiLU->Conv3d(3,1,1)."""
    identity = x
    for i, conv_idx in ((1, 2), (2, 3), (3, 3), (4, 3)):
        q = p.sub(f"conv{i}")
        x = F.silu(_gn(q.sub("0"), x, 1e-5))
        x = F.conv3d(x, q[f"{conv_idx}.weight"], q[f"{conv_idx}.bias"], padding=(1, 0, 0))
    return x + identity


def res_block(p: _SD, x, emb, batch_size, temporal_conv):
    """openaimodel3d.py:223-254 (no up/down, use_scale_shift_norm=False)."""
    h = F.silu(_gn(p.sub("in_layers.0"), x, 1e-5))
    h = F.conv2d(h, p["in_layers.2.weight"], p["in_layers.2.bias"], padding=1)
    emb_out = F.linear(F.silu(emb), p["emb_layers.1.weight"], p["emb_layers.1.bias"]).type(h.dtype)
    h = h + emb_out[:, :, None, None]
    h = F.silu(_gn(p.sub("out_layers.0"), h, 1e-5))
    h = F.conv2d(h, p["out_layers.3.weight"], p["out_layers.3.bias"], padding=1)
    if p.has("skip_connection.weight"):
        x = F.conv2d(x, p["skip_connection.weight"], p["skip_connection.bias"])
    h = x + h
    if temporal_conv:
        bt, c, hh, ww = h.shape
        h5 = h.reshape(batch_size, bt // batch_size, c, hh, ww).permute(0, 2, 1, 3, 4)
        h5 = temporal_conv_block(p.sub("temopral_conv"), h5)
        h = h5.permute(0, 2, 1, 3, 4).reshape(bt, c, hh, ww)
    return h


def unet_layout(cfg):
    """Module census of UNetModel.__init__ (openaimodel3d.py:433-670): list of blocks, each a list of
    (kind, name_suffix, heads)."""
    mc = cfg["model_channels"]
    mult = cfg["channel_mult"]
    nrb = cfg["num_res_blocks"]
    ar = cfg["attention_resolutions"]
    hd = cfg["num_head_channels"]
    inp = [[("conv_in", "0", 0)]]
    ch, ds = mc, 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            layers = [("res", "0", 0)]
            ch = m * mc
            if ds in ar:
                layers += [("st", "1", ch // hd), ("tt", "2", ch // hd)]
            inp.append(layers)
        if level != len(mult) - 1:
            inp.append([("down", "0", 0)])
            ds *= 2
    mid = [("res", "0", 0), ("st", "1", ch // hd), ("tt", "2", ch // hd), ("res", "3", 0)]
    out = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            layers = [("res", "0", 0)]
            ch = m * mc
            if ds in ar:
                layers += [("st", "1", ch // hd), ("tt", "2", ch // hd)]
            if level and i == nrb:
                layers.append(("up", str(len(layers)), 0))
                ds //= 2
            out.append(layers)
    return inp, mid, out


def unet_forward(sd, cfg, x, timesteps, context, fps=16, timestep_cond=None, motion_cond=None, dtype=None):
    """openaimodel3d.py:672-740.  sd: state dict (reference key names); cfg: UNetModel kwargs."""
    p = _SD(sd)
    dtype = dtype or sd["time_embed.0.weight"].dtype
    mc = cfg["model_channels"]
    temporal_conv = cfg.get("temporal_conv", False)
    t_emb = timestep_embedding(timesteps, mc).to(dtype)
    cond = 0.0
    if timestep_cond is not None:
        cond = F.linear(timestep_cond.to(dtype), sd["time_cond_proj.weight"])
    if motion_cond is not None:
        m = F.linear(motion_cond.to(dtype), sd["motion_cond_proj.weight"])
        cond = F.linear(torch.cat([cond, m], dim=1), sd["combine_proj.weight"])

    def mlp(name, v):
        v = F.silu(_linear(p.sub(f"{name}.0"), v))
        return _linear(p.sub(f"{name}.2"), v)
    emb = mlp("time_embed", t_emb + cond)
    if cfg.get("fps_cond", False):
        if isinstance(fps, int):
            fps = torch.full_like(timesteps, fps)
        emb = emb + mlp("fps_embedding", timestep_embedding(fps, mc).to(dtype))
    b, _, t, _, _ = x.shape
    context = context.to(dtype).repeat_interleave(repeats=t, dim=0)
    emb = emb.repeat_interleave(repeats=t, dim=0)
    h = x.permute(0, 2, 1, 3, 4).reshape(b * t, x.shape[1], x.shape[3], x.shape[4]).to(dtype)

    def run(prefix, layers, h):
        for kind, idx, heads in layers:
            q = p.sub(f"{prefix}.{idx}")
            if kind == "conv_in":
                h = F.conv2d(h, q["weight"], q["bias"], padding=1)
            elif kind == "res":
                h = res_block(q, h, emb, b, temporal_conv)
            elif kind == "st":
                h = spatial_transformer(q, h, context, heads)
            elif kind == "tt":
                bt, c, hh, ww = h.shape
                h5 = h.reshape(b, bt // b, c, hh, ww).permute(0, 2, 1, 3, 4)
                h5 = temporal_transformer(q, h5, heads)
                h = h5.permute(0, 2, 1, 3, 4).reshape(bt, c, hh, ww)
            elif kind == "down":
                h = F.conv2d(h, q["op.weight"], q["op.bias"], stride=2, padding=1)
            elif kind == "up":
                h = F.interpolate(h, scale_factor=2, mode="nearest")
                h = F.conv2d(h, q["conv.weight"], q["conv.bias"], padding=1)
        return h

    inp, mid, out = unet_layout(cfg)
    hs = []
    for i, layers in enumerate(inp):
        h = run(f"input_blocks.{i}", layers, h)
        if i == 0 and cfg.get("addition_attention", False):
            h = run("init_attn", [("tt", "0", 8)], h)
        hs.append(h)
    h = run("middle_block", mid, h)
    for i, layers in enumerate(out):
        h = torch.cat([h, hs.pop()], dim=1)
        h = run(f"output_blocks.{i}", layers, h)
    h = F.silu(_gn(p.sub("out.0"), h, 1e-5))
    y = F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)
    return y.reshape(b, t, y.shape[1], y.shape[2], y.shape[3]).permute(0, 2, 1, 3, 4)
