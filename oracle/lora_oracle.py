"""ORACLE (test infrastructure): the LoRA-injected layers of utils/lora.py:19-230 restated as plain fp32 torch, with
autograd for the backward (dx, d lora_up, d lora_down) — `y = base(x) + dropout(up(down(x))) * scale`, dropout off."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def lora_layer(kind, x, w, bias, up, down, scale):
    """kind: linear | conv2d (3x3, pad 1) | conv3d ((3,1,1), pad (1,0,0)); layouts as in the reference (NCHW / NCDHW)."""
    if kind == "linear":
        return F.linear(x, w, bias) + F.linear(F.linear(x, down), up) * scale
    if kind == "conv2d":
        return F.conv2d(x, w, bias, padding=1) + F.conv2d(F.conv2d(x, down, padding=1), up) * scale
    return F.conv3d(x, w, bias, padding=(1, 0, 0)) + F.conv3d(F.conv3d(x, down, padding=(1, 0, 0)), up) * scale


def lora_layer_grads(kind, x, w, bias, up, down, scale, dy):
    x = x.clone().requires_grad_(True)
    up = up.clone().requires_grad_(True)
    down = down.clone().requires_grad_(True)
    y = lora_layer(kind, x, w, bias, up, down, scale)
    y.backward(dy)
    return y.detach(), x.grad, up.grad, down.grad
