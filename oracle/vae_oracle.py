"""ORACLE (test infrastructure): plain-torch restatement of the reference KL-VAE decode path —
`LatentDiffusion.decode_first_stage_2DAE` (lvdm/models/ddpm3d.py:666-679) -> `AutoencoderKL.decode`
(lvdm/models/autoencoder.py:110-113) -> `Decoder.forward` (lvdm/modules/networks/ae_modules.py:602-641).
State-dict keys are the reference's (`post_quant_conv.*`, `decoder.*`).  Pinned by
oracle/make_goldens.py against the unmodified reference modules (tests/golden/vae_*.pt)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _norm(sd, name, x):
    # ae_modules.py:16-19 Normalize: GroupNorm(32, C, eps=1e-6, affine)
    return F.group_norm(x, 32, sd[f"{name}.weight"], sd[f"{name}.bias"], 1e-6)


def _swish(x):  # ae_modules.py:11-13
    return x * torch.sigmoid(x)


def resnet_block(sd, name, x):
    """ae_modules.py:183-203 with temb=None."""
    h = F.conv2d(_swish(_norm(sd, f"{name}.norm1", x)), sd[f"{name}.conv1.weight"], sd[f"{name}.conv1.bias"], padding=1)
    h = F.conv2d(_swish(_norm(sd, f"{name}.norm2", h)), sd[f"{name}.conv2.weight"], sd[f"{name}.conv2.bias"], padding=1)
    if f"{name}.nin_shortcut.weight" in sd:
        x = F.conv2d(x, sd[f"{name}.nin_shortcut.weight"], sd[f"{name}.nin_shortcut.bias"])
    return x + h


def attn_block(sd, name, x):
    """ae_modules.py:48-73: single-head attention over h*w tokens, scale c^-0.5."""
    h_ = _norm(sd, f"{name}.norm", x)
    q = F.conv2d(h_, sd[f"{name}.q.weight"], sd[f"{name}.q.bias"])
    k = F.conv2d(h_, sd[f"{name}.k.weight"], sd[f"{name}.k.bias"])
    v = F.conv2d(h_, sd[f"{name}.v.weight"], sd[f"{name}.v.bias"])
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    h_ = F.conv2d(h_, sd[f"{name}.proj_out.weight"], sd[f"{name}.proj_out.bias"])
    return x + h_


def decoder_forward(sd, ddconfig, z, prefix="decoder"):
    """ae_modules.py:602-641 (attn_resolutions=[], give_pre_end=False, tanh_out=False)."""
    nres = len(ddconfig["ch_mult"])
    nrb = ddconfig["num_res_blocks"]
    h = F.conv2d(z, sd[f"{prefix}.conv_in.weight"], sd[f"{prefix}.conv_in.bias"], padding=1)
    h = resnet_block(sd, f"{prefix}.mid.block_1", h)
    h = attn_block(sd, f"{prefix}.mid.attn_1", h)
    h = resnet_block(sd, f"{prefix}.mid.block_2", h)
    for i_level in reversed(range(nres)):
        for i_block in range(nrb + 1):
            h = resnet_block(sd, f"{prefix}.up.{i_level}.block.{i_block}", h)
        if i_level != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[f"{prefix}.up.{i_level}.upsample.conv.weight"], sd[f"{prefix}.up.{i_level}.upsample.conv.bias"], padding=1)
    h = _swish(_norm(sd, f"{prefix}.norm_out", h))
    return F.conv2d(h, sd[f"{prefix}.conv_out.weight"], sd[f"{prefix}.conv_out.bias"], padding=1)


def decode_first_stage_2dae(sd, ddconfig, z, scale_factor=0.18215):
    """ddpm3d.py:666-679: per-frame decode of z [b, c, t, h, w] -> [b, 3, t, 8h, 8w]."""
    z = 1.0 / scale_factor * z
    frames = []
    for i in range(z.shape[2]):
        zi = F.conv2d(z[:, :, i], sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])  # autoencoder.py:111
        frames.append(decoder_forward(sd, ddconfig, zi).unsqueeze(2))
    return torch.cat(frames, dim=2)


# ------------------------------------------------------------------------------------------ encode path
def encoder_forward(sd, ddconfig, x, prefix="encoder"):
    """ae_modules.py:470-503 `Encoder.forward` (attn_resolutions=[], temb=None).  Downsample (ae_modules.py:87-105):
    zero-pad right/bottom by one, then a stride-2 3x3 conv without padding."""
    nres = len(ddconfig["ch_mult"])
    nrb = ddconfig["num_res_blocks"]
    h = F.conv2d(x, sd[f"{prefix}.conv_in.weight"], sd[f"{prefix}.conv_in.bias"], padding=1)
    for i_level in range(nres):
        for i_block in range(nrb):
            h = resnet_block(sd, f"{prefix}.down.{i_level}.block.{i_block}", h)
        if i_level != nres - 1:
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)
            h = F.conv2d(h, sd[f"{prefix}.down.{i_level}.downsample.conv.weight"],
                         sd[f"{prefix}.down.{i_level}.downsample.conv.bias"], stride=2)
    h = resnet_block(sd, f"{prefix}.mid.block_1", h)
    h = attn_block(sd, f"{prefix}.mid.attn_1", h)
    h = resnet_block(sd, f"{prefix}.mid.block_2", h)
    h = _swish(_norm(sd, f"{prefix}.norm_out", h))
    return F.conv2d(h, sd[f"{prefix}.conv_out.weight"], sd[f"{prefix}.conv_out.bias"], padding=1)


def encode_moments(sd, ddconfig, x):
    """autoencoder.py:103-107: moments = quant_conv(encoder(x)); x [n, 3, H, W] -> [n, 2*embed_dim, H/8, W/8]."""
    return F.conv2d(encoder_forward(sd, ddconfig, x), sd["quant_conv.weight"], sd["quant_conv.bias"])


def encode_first_stage(sd, ddconfig, x, noise=None, scale_factor=0.18215):
    """ddpm3d.py:558-584: x [b, 3, t, H, W] -> scale_factor * (mean + std * noise), [b, c, t, H/8, W/8].
    distributions.py:24-42: logvar clamped to [-30, 20], std = exp(logvar / 2); noise=None -> posterior mode (mean)."""
    b, _, t = x.shape[:3]
    frames = x.permute(0, 2, 1, 3, 4).reshape(b * t, *x.shape[1:2], *x.shape[3:])
    mom = encode_moments(sd, ddconfig, frames)
    mean, logvar = torch.chunk(mom, 2, dim=1)
    z = mean
    if noise is not None:
        z = mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * noise
    z = scale_factor * z
    return z.reshape(b, t, *z.shape[1:]).permute(0, 2, 1, 3, 4)
