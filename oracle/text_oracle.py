"""ORACLE (test infrastructure; pinned on transformers.CLIPTextModel, an independent implementation of the same tower, in
tests/test_text_encoder.py — open_clip itself is absent): the OpenCLIP ViT-H-14 text tower as FrozenOpenCLIPEmbedder drives it
(lvdm/modules/encoders/condition.py:257-283).  open_clip is not installed here; this restates its published text
transformer (open_clip/transformer.py: ResidualAttentionBlock = x + attn(ln_1(x)), x + mlp(ln_2(x)); nn.MultiheadAttention
with an additive -inf upper-triangular mask; mlp = c_fc -> GELU(erf) -> c_proj) under open_clip's key names."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def text_forward(sd, tokens, heads=16, layer_idx=1):
    """sd: state dict with `model.` keys; tokens int64 [B, n]; layer_idx 1 = penultimate (skip the last block)."""
    x = sd["model.token_embedding.weight"][tokens] + sd["model.positional_embedding"][: tokens.shape[1]]
    n_blocks = len({k.split(".")[3] for k in sd if k.startswith("model.transformer.resblocks.")})
    b, n, w = x.shape
    d = w // heads
    mask = torch.full((n, n), float("-inf")).triu_(1)
    for i in range(n_blocks - layer_idx):
        p = f"model.transformer.resblocks.{i}."
        y = F.layer_norm(x, (w,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-5)
        qkv = F.linear(y, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"])
        q, k, v = (t.reshape(b, n, heads, d).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
        att = torch.softmax(q @ k.transpose(-1, -2) * d ** -0.5 + mask, dim=-1) @ v
        x = x + F.linear(att.transpose(1, 2).reshape(b, n, w), sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])
        y = F.layer_norm(x, (w,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], 1e-5)
        y = F.gelu(F.linear(y, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"]))
        x = x + F.linear(y, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
    return F.layer_norm(x, (w,), sd["model.ln_final.weight"], sd["model.ln_final.bias"], 1e-5)
