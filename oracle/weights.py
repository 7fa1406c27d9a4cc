"""Deterministic weights for parity tests (test infrastructure).

No checkpoint exists offline and the reference's default init leaves every residual branch at zero
(`zero_module`: openaimodel3d.py:183,299-300,669; attention.py:366-370,464-468), so goldens must use
re-randomised parameters.  `seeded_state_dict` depends only on (sorted key, shape), so the reference
module in the authoring container and the B200 module on the GPU box get bit-identical fp32 weights
(torch's CPU generator is platform-independent for a fixed torch version)."""
from __future__ import annotations

import math

import torch


def seeded_state_dict(template: dict, seed: int) -> dict:
    """template: name -> tensor (only shapes/dtypes are used). Returns name -> fp32 tensor."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in sorted(template.keys()):
        t = template[name]
        shape = tuple(t.shape)
        if not t.dtype.is_floating_point:
            out[name] = t.clone()
            continue
        if len(shape) >= 2:                       # Linear / Conv weights
            fan_in = math.prod(shape[1:])
            v = torch.randn(shape, generator=g) * (0.8 / math.sqrt(fan_in))
        elif name.endswith("weight"):              # norm gains
            v = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:                                      # biases
            v = 0.05 * torch.randn(shape, generator=g)
        out[name] = v
    return out


def vae_state_dict(module_sd: dict, seed: int) -> dict:
    """AutoencoderKL weights for the parity tests, identical to what oracle/make_goldens.py loaded into the reference:
    the decode side (`decoder.*`, `post_quant_conv.*`) from `seed`, the encode side (`encoder.*`, `quant_conv.*`)
    from `seed + 100`, each generated over its own sorted key set."""
    dec = {k: v for k, v in module_sd.items() if k.startswith(("decoder.", "post_quant_conv."))}
    enc = {k: v for k, v in module_sd.items() if k.startswith(("encoder.", "quant_conv."))}
    out = seeded_state_dict(dec, seed)
    out.update(seeded_state_dict(enc, seed + 100))
    return out
