"""LoRA checkpoints for the B200 UNet: the reference's `unet_lora.pt` wire format and its merge.

T2V-Turbo ships its distilled weights as a LoRA over VideoCrafter2 (`app.py:244-265`): the reference injects
`LoraInjectedLinear / Conv2d / Conv3d` wrappers into every `nn.Linear / nn.Conv2d / nn.Conv3d` below the `UNetModel`
(`utils/lora.py:387-486`, traversal `_find_modules_v2` :263-307), loads `unet_lora.pt` — a FLAT LIST
`[up_0, down_0, up_1, down_1, ...]` in traversal order (`save_lora_weight` :581-594, loader :466-477) — and then
collapses it, `W += alpha * (up.flatten(1) @ down.flatten(1)).reshape(W.shape)` (`collapse_lora` :793-830).
Inference therefore only ever needs the merged weights: this module reproduces the traversal order on the B200
`UNetModel` (whose module tree mirrors the reference's) and merges the list straight into the base parameters —
load-time fp32 arithmetic, after which `UNetModel.pack()` builds the bf16 UMMA operands as usual.
"""
from __future__ import annotations

import torch
import torch.nn as nn

_SEARCH = (nn.Linear, nn.Conv2d, nn.Conv3d)


def lora_target_layers(unet: nn.Module):
    """[(qualified name, module)] in the order of the reference's flat list: every module below `unet` whose class is
    exactly nn.Linear / nn.Conv2d / nn.Conv3d (nn.Conv1d is not wrapped: utils/lora.py:404-452), in
    `named_modules()` order (`_find_modules_v2` with ancestor class {"UNetModel"})."""
    return [(name, m) for name, m in unet.named_modules() if m.__class__ in _SEARCH]


def lora_rank(m: nn.Module, r: int = 64) -> int:
    """LoraInjected*.__init__ (utils/lora.py:19-230): r is clamped to min(r, in, out)."""
    cin = m.in_features if isinstance(m, nn.Linear) else m.in_channels
    cout = m.out_features if isinstance(m, nn.Linear) else m.out_channels
    return min(r, cin, cout)


def lora_shapes(unet: nn.Module, r: int = 64):
    """Expected shapes of the flat list: [(up shape, down shape)] per target layer.  Linear: up [out, r], down [r, in];
    conv: down carries the base kernel size [r, in, *k], up is 1x1 [out, r, 1, ...] (utils/lora.py:69-230)."""
    out = []
    for _, m in lora_target_layers(unet):
        rr = lora_rank(m, r)
        if isinstance(m, nn.Linear):
            out.append(((m.out_features, rr), (rr, m.in_features)))
        else:
            k = tuple(m.kernel_size)
            out.append(((m.out_channels, rr) + (1,) * len(k), (rr, m.in_channels) + k))
    return out


@torch.no_grad()
def merge_lora(unet: nn.Module, loras, alpha: float = 1.0, r: int = 64) -> int:
    """Merge a reference `unet_lora.pt` (path or the loaded flat list) into the UNet's base weights — the net effect
    of app.py:250-265 (inject + load + collapse_lora + monkeypatch_remove_lora).  Returns the number of merged layers."""
    if isinstance(loras, (str, bytes)) or hasattr(loras, "__fspath__"):
        loras = torch.load(loras, map_location="cpu", weights_only=True)
    layers = lora_target_layers(unet)
    if len(loras) != 2 * len(layers):
        raise ValueError(f"LoRA list has {len(loras)} tensors, the UNet has {len(layers)} target layers (expected {2 * len(layers)})")
    shapes = lora_shapes(unet, r)
    for i, ((name, m), (us, ds)) in enumerate(zip(layers, shapes)):
        up, down = loras[2 * i], loras[2 * i + 1]
        if tuple(up.shape) != us or tuple(down.shape) != ds:
            raise ValueError(f"LoRA tensors for {name}: got {tuple(up.shape)} / {tuple(down.shape)}, expected {us} / {ds}")
        w = m.weight
        delta = up.to(device=w.device, dtype=torch.float32).flatten(1) @ down.to(device=w.device, dtype=torch.float32).flatten(1)
        w.data = (w.data.float() + alpha * delta.reshape(w.shape)).to(w.dtype)
    if hasattr(unet, "invalidate_packed"):
        unet.invalidate_packed()
    return len(layers)
