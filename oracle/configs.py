"""Shared parity configurations (test infrastructure): model kwargs, seeds and seeded inputs used by
oracle/make_goldens.py (reference side, authoring container) and tests/ (oracle + B200 side)."""
from __future__ import annotations

import torch

from t2v_turbo_b200.configs import VC2_UNET, VC2_VAE_DDCONFIG  # noqa: E402,F401

UNET_CONFIGS = {
    # two levels, 64/128 channels, every block type; seconds on CPU
    "small": dict(cfg={**VC2_UNET, "model_channels": 64, "attention_resolutions": [2, 1], "num_res_blocks": 1,
                       "channel_mult": [1, 2], "context_dim": 128, "temporal_length": 4},
                  x_shape=(1, 4, 4, 8, 8), ctx_len=77, weight_seed=7, input_seed=1234, timesteps=[999, 519]),
    # all four levels of the VC2 topology at 128 base channels (head counts 2/4/8/8, straddling GN groups)
    "mid": dict(cfg={**VC2_UNET, "model_channels": 128, "temporal_length": 8},
                x_shape=(2, 4, 8, 16, 16), ctx_len=77, weight_seed=8, input_seed=1235, timesteps=[759]),
    # BASELINE.json config 1: full VC2 UNet, 1x4x16x40x64 latent, 77x1024 text embedding
    "full": dict(cfg=dict(VC2_UNET), x_shape=(1, 4, 16, 40, 64), ctx_len=77, weight_seed=9, input_seed=1234,
                 timesteps=[999]),
    # the same model at the later sampling timesteps of the 4-step schedule, and at batch 2 with a different timestep
    # per sample (round 2; separate fixtures so the round-1 golden stays untouched)
    "full_t": dict(cfg=dict(VC2_UNET), x_shape=(1, 4, 16, 40, 64), ctx_len=77, weight_seed=9, input_seed=1236,
                   timesteps=[759, 279]),
    "full_b2": dict(cfg=dict(VC2_UNET), x_shape=(2, 4, 16, 40, 64), ctx_len=77, weight_seed=9, input_seed=1237,
                    timesteps=[(519, 279)]),
    # v2 models (BASELINE config 3): motion_cond_proj_dim=256 adds motion_cond_proj / combine_proj (openaimodel3d.py:690-697)
    "small_motion": dict(cfg={**VC2_UNET, "model_channels": 64, "attention_resolutions": [2, 1], "num_res_blocks": 1,
                              "channel_mult": [1, 2], "context_dim": 128, "temporal_length": 4, "motion_cond_proj_dim": 256},
                         x_shape=(2, 4, 4, 8, 8), ctx_len=77, weight_seed=17, input_seed=1238, timesteps=[(999, 519)],
                         motion_gs=(0.05, 0.0)),
}

VAE_CONFIGS = {
    "small": dict(ddconfig={**VC2_VAE_DDCONFIG, "ch": 64, "ch_mult": [1, 2, 2], "num_res_blocks": 1, "resolution": 64},
                  embed_dim=4, z_shape=(1, 4, 4, 16, 16), weight_seed=21, input_seed=22),
    "full": dict(ddconfig=dict(VC2_VAE_DDCONFIG), embed_dim=4, z_shape=(1, 4, 1, 40, 64), weight_seed=23, input_seed=24),
}


def unet_inputs(spec, timestep):
    """Seeded inputs: x ~ N(0,1), ctx ~ N(0,1), w-embedding of guidance 7.5, fps 16."""
    from oracle.unet_oracle import guidance_scale_embedding
    g = torch.Generator().manual_seed(spec["input_seed"])
    b = spec["x_shape"][0]
    x = torch.randn(spec["x_shape"], generator=g)
    ctx = torch.randn(b, spec["ctx_len"], spec["cfg"]["context_dim"], generator=g)
    w_emb = guidance_scale_embedding(torch.tensor(7.5).repeat(b), 256)
    ts = torch.tensor(timestep, dtype=torch.long) if isinstance(timestep, (tuple, list)) else torch.full((b,), timestep, dtype=torch.long)
    out = dict(x=x, context=ctx, timesteps=ts, fps=16, timestep_cond=w_emb)
    if "motion_gs" in spec:   # pipeline:197-204: the same sin/cos embedding of the motion guidance scale
        out["motion_cond"] = guidance_scale_embedding(torch.tensor(spec["motion_gs"], dtype=torch.float32), 256)
    return out


def student_loras(shapes, seed=4242):
    """Seeded LoRA weights for the student-gradient fixture, in the reference's flat wire order [up_0, down_0, ...]: both
    factors non-zero (the reference initialises up = 0, which would zero every lora_down gradient)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(0, len(shapes), 2):
        up_s, down_s = shapes[i], shapes[i + 1]
        r = up_s[1]
        out.append(torch.randn(up_s, generator=g) * (0.5 / r ** 0.5))
        fan = 1
        for d in down_s[1:]:
            fan *= d
        out.append(torch.randn(down_s, generator=g) * (0.5 / fan ** 0.5))
    return out
