"""ModelScope path (BASELINE config 5): `UNet3DConditionModel` + `T2VTurboMSPipeline` on the B200 kernels.

The reference's ModelScope variant (model_scope/unet_3d_condition.py:329-504, unet_3d_blocks.py:268-875,
pipeline/t2v_turbo_ms_pipeline.py:132-221) wires diffusers building blocks — ResnetBlock2D, TemporalConvLayer,
Transformer2DModel, TransformerTemporalModel, Downsample2D / Upsample2D — into a UNet whose dataflow is the SAME as the
VideoCrafter2 UNet this package already runs: per layer  resnet -> temporal conv -> spatial transformer (cross-attention on
the frame-repeated text states) -> temporal transformer (double self-attention);  channels (320, 640, 1280, 1280), two
layers per level, attention on the three upper levels, a temporal transformer of 8 x 64 right after conv_in
(`transformer_in` == VC2's `init_attn`), the same sinusoidal embedding ([cos | sin], freq / half), the same
`time_embedding.cond_proj` for the guidance-scale embedding, GroupNorm eps 1e-5 in the resnets / temporal convs / output
norm and 1e-6 in the transformers.  What differs is the NAMING (diffusers state-dict keys), the absence of the fps
embedding, `nn.Linear` instead of `Conv1d(k=1)` for `transformer_in.proj_in / proj_out`, fp16 as the serving dtype and
32 x 32 latents.  So this adapter is a weight-NAME map over the same kernels:

    diffusers key                                         B200 `UNetModel` key
    conv_in                                               input_blocks.0.0
    transformer_in.*                                      init_attn.0.*           (proj_in / proj_out: [o, i] -> [o, i, 1])
    down_blocks.L.resnets.M.{norm1,conv1,time_emb_proj,   input_blocks.(1+3L+M).0.{in_layers.0,in_layers.2,emb_layers.1,
        norm2,conv2,conv_shortcut}                            out_layers.0,out_layers.3,skip_connection}
    down_blocks.L.temp_convs.M.convK.*                    input_blocks.(1+3L+M).0.temopral_conv.convK.*
    down_blocks.L.attentions.M.* / temp_attentions.M.*    input_blocks.(1+3L+M).1.* / .2.*
    down_blocks.L.downsamplers.0.conv                     input_blocks.(3+3L).0.op
    mid_block.{resnets.0,attentions.0,temp_attentions.0,  middle_block.{0,1,2,3} (+ temp_convs -> temopral_conv)
        resnets.1}
    up_blocks.U.resnets.M ... / upsamplers.0.conv         output_blocks.(3U+M).0 ... / output_blocks.(3U+2).{1|3}.conv
    time_embedding.{linear_1,linear_2,cond_proj}          time_embed.{0,2} / time_cond_proj
    conv_norm_out / conv_out                              out.0 / out.2

and the same for the SD KL-VAE the pipeline decodes with (diffusers `AutoencoderKL` names -> lvdm names).  Compute is bf16
inside (fp32 accumulation) with fp16 / fp32 tensors accepted and returned at the boundary.

PARITY UNPINNED for the diffusers arithmetic: diffusers is not installed here, so the oracle for this path
(oracle/ms_oracle.py) is a restatement of diffusers 0.30.0's blocks wired as in model_scope/unet_3d_blocks.py, not an
execution of them; the tests check this adapter against that restatement and the key map against the module census.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn as nn

from .configs import VC2_UNET, VC2_VAE_DDCONFIG
from .pipeline import _GraphedCall
from .unet import UNetModel
from .vae import AutoencoderKL

# the B200 UNetModel kwargs that reproduce UNet3DConditionModel's defaults (unet_3d_condition.py:87-107)
MS_UNET = {**VC2_UNET, "fps_cond": False, "addition_attention": True, "time_cond_proj_dim": 256, "temporal_length": 16}

_RESNET = {"norm1": "in_layers.0", "conv1": "in_layers.2", "time_emb_proj": "emb_layers.1", "norm2": "out_layers.0",
           "conv2": "out_layers.3", "conv_shortcut": "skip_connection"}


def ms_unet_key_map(block_out_channels=(320, 640, 1280, 1280), layers_per_block=2) -> dict:
    """{diffusers key prefix: B200 key prefix} for every parameterised sub-module of UNet3DConditionModel."""
    n_levels = len(block_out_channels)
    m = {"conv_in": "input_blocks.0.0", "transformer_in": "init_attn.0", "time_embedding.linear_1": "time_embed.0",
         "time_embedding.linear_2": "time_embed.2", "time_embedding.cond_proj": "time_cond_proj", "conv_norm_out": "out.0",
         "conv_out": "out.2"}

    def layer(src, dst, attn):
        for a, b in _RESNET.items():
            m[f"{src[0]}.{a}"] = f"{dst}.0.{b}"
        m[src[1]] = f"{dst}.0.temopral_conv"
        if attn:
            m[src[2]] = f"{dst}.1"
            m[src[3]] = f"{dst}.2"
    per = layers_per_block + 1
    for lvl in range(n_levels):
        attn = lvl < n_levels - 1                      # CrossAttnDownBlock3D x3, DownBlock3D
        for j in range(layers_per_block):
            d = f"down_blocks.{lvl}"
            layer((f"{d}.resnets.{j}", f"{d}.temp_convs.{j}", f"{d}.attentions.{j}", f"{d}.temp_attentions.{j}"),
                  f"input_blocks.{1 + per * lvl + j}", attn)
        if lvl < n_levels - 1:
            m[f"down_blocks.{lvl}.downsamplers.0.conv"] = f"input_blocks.{per * (lvl + 1)}.0.op"
    for a, b in _RESNET.items():
        m[f"mid_block.resnets.0.{a}"] = f"middle_block.0.{b}"
        m[f"mid_block.resnets.1.{a}"] = f"middle_block.3.{b}"
    m["mid_block.temp_convs.0"] = "middle_block.0.temopral_conv"
    m["mid_block.temp_convs.1"] = "middle_block.3.temopral_conv"
    m["mid_block.attentions.0"] = "middle_block.1"
    m["mid_block.temp_attentions.0"] = "middle_block.2"
    for u in range(n_levels):
        attn = u > 0                                    # UpBlock3D, CrossAttnUpBlock3D x3
        for j in range(per):
            s = f"up_blocks.{u}"
            layer((f"{s}.resnets.{j}", f"{s}.temp_convs.{j}", f"{s}.attentions.{j}", f"{s}.temp_attentions.{j}"),
                  f"output_blocks.{per * u + j}", attn)
        if u < n_levels - 1:
            m[f"up_blocks.{u}.upsamplers.0.conv"] = f"output_blocks.{per * u + per - 1}.{3 if attn else 1}.conv"
    return m


def _remap(sd: dict, prefix_map: dict, what: str) -> dict:
    out = {}
    prefixes = sorted(prefix_map, key=len, reverse=True)
    for k, v in sd.items():
        for p in prefixes:
            if k == p or k.startswith(p + "."):
                out[prefix_map[p] + k[len(p):]] = v
                break
        else:
            raise KeyError(f"{what}: no mapping for key {k!r}")
    return out


def convert_ms_unet_state_dict(sd: dict, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2) -> dict:
    """diffusers `UNet3DConditionModel.state_dict()` -> B200 `UNetModel(**MS_UNET)` state dict."""
    out = _remap(sd, ms_unet_key_map(block_out_channels, layers_per_block), "UNet3DConditionModel")
    for k in ("init_attn.0.proj_in.weight", "init_attn.0.proj_out.weight"):   # nn.Linear -> Conv1d(kernel 1)
        if out[k].dim() == 2:
            out[k] = out[k].unsqueeze(-1)
    return out


def diffusers_vae_key_map(n_levels=4, num_res_blocks=2) -> dict:
    """{diffusers AutoencoderKL key prefix: lvdm AutoencoderKL key prefix} (autoencoder_kl.py / vae.py of diffusers 0.30 vs
    lvdm/modules/networks/ae_modules.py): lvdm indexes `up` by resolution level (up.3 runs first), diffusers by order."""
    res = {"norm1": "norm1", "conv1": "conv1", "norm2": "norm2", "conv2": "conv2", "conv_shortcut": "nin_shortcut"}
    attn = {"group_norm": "norm", "to_q": "q", "to_k": "k", "to_v": "v", "to_out.0": "proj_out"}
    m = {"quant_conv": "quant_conv", "post_quant_conv": "post_quant_conv"}
    for side in ("encoder", "decoder"):
        m[f"{side}.conv_in"] = f"{side}.conv_in"
        m[f"{side}.conv_norm_out"] = f"{side}.norm_out"
        m[f"{side}.conv_out"] = f"{side}.conv_out"
        for i, blk in ((0, "block_1"), (1, "block_2")):
            for a, b in res.items():
                m[f"{side}.mid_block.resnets.{i}.{a}"] = f"{side}.mid.{blk}.{b}"
        for a, b in attn.items():
            m[f"{side}.mid_block.attentions.0.{a}"] = f"{side}.mid.attn_1.{b}"
    for lvl in range(n_levels):
        for j in range(num_res_blocks):
            for a, b in res.items():
                m[f"encoder.down_blocks.{lvl}.resnets.{j}.{a}"] = f"encoder.down.{lvl}.block.{j}.{b}"
        m[f"encoder.down_blocks.{lvl}.downsamplers.0.conv"] = f"encoder.down.{lvl}.downsample.conv"
        for j in range(num_res_blocks + 1):
            for a, b in res.items():
                m[f"decoder.up_blocks.{lvl}.resnets.{j}.{a}"] = f"decoder.up.{n_levels - 1 - lvl}.block.{j}.{b}"
        m[f"decoder.up_blocks.{lvl}.upsamplers.0.conv"] = f"decoder.up.{n_levels - 1 - lvl}.upsample.conv"
    return m


def convert_diffusers_vae_state_dict(sd: dict) -> dict:
    out = _remap(sd, diffusers_vae_key_map(), "AutoencoderKL")
    for k, v in out.items():   # diffusers keeps the mid-block attention projections as nn.Linear, lvdm as 1x1 convs
        if ".mid.attn_1." in k and k.endswith("weight") and v.dim() == 2:
            out[k] = v[:, :, None, None]
    return out


class UNet3DConditionModel(nn.Module):
    """Drop-in for model_scope/unet_3d_condition.py:UNet3DConditionModel on B200: same `forward(sample, timestep,
    encoder_hidden_states, timestep_cond=...)` -> `.sample` ([B, 4, F, h, w] in and out), `config.in_channels`, and
    `load_state_dict` of a diffusers-keyed checkpoint (converted on the fly; B200-keyed dicts load as they are)."""

    def __init__(self, sample_size=None, in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                 norm_num_groups=32, norm_eps=1e-5, time_cond_proj_dim=256, cross_attention_dim=1024, attention_head_dim=64, **ignored):
        super().__init__()
        boc = tuple(block_out_channels)
        if norm_num_groups != 32 or attention_head_dim != 64 or any(c % boc[0] or c % 64 for c in boc) or norm_eps != 1e-5:
            raise NotImplementedError("UNet3DConditionModel(B200): 32 norm groups, eps 1e-5, 64-wide heads, channel multiples of the first level")
        self.config = SimpleNamespace(sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
                                      block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                                      time_cond_proj_dim=time_cond_proj_dim, cross_attention_dim=cross_attention_dim,
                                      attention_head_dim=attention_head_dim, norm_eps=norm_eps)
        # CrossAttn blocks on every level but the last (unet_3d_condition.py:91-97) == attention at downsample rates 1, 2, 4, ...
        self.model = UNetModel(**{**MS_UNET, "in_channels": in_channels, "out_channels": out_channels, "context_dim": cross_attention_dim,
                                  "time_cond_proj_dim": time_cond_proj_dim, "model_channels": boc[0],
                                  "channel_mult": [c // boc[0] for c in boc], "num_res_blocks": layers_per_block,
                                  "attention_resolutions": [2 ** i for i in range(len(boc) - 1)]})

    @property
    def dtype(self):
        return next(self.model.parameters()).dtype

    def load_state_dict(self, sd, strict=True):
        if any(k.startswith(("down_blocks.", "up_blocks.", "mid_block.")) for k in sd):
            sd = convert_ms_unet_state_dict(sd, self.config.block_out_channels, self.config.layers_per_block)
        elif any(k.startswith("model.") for k in sd):
            sd = {k[len("model."):]: v for k, v in sd.items()}
        return self.model.load_state_dict(sd, strict=strict)

    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, timestep_cond=None, attention_mask=None,
                cross_attention_kwargs=None, return_dict=True, **kwargs):
        if class_labels is not None or attention_mask is not None:
            raise NotImplementedError("UNet3DConditionModel(B200): class_labels / attention_mask are not used by T2V-Turbo")
        ts = timestep if torch.is_tensor(timestep) else torch.tensor([timestep], device=sample.device)
        ts = ts.reshape(-1).expand(sample.shape[0]) if ts.numel() == 1 else ts
        self.model.dtype = torch.bfloat16 if sample.dtype in (torch.float16, torch.bfloat16) else torch.float32
        out = self.model(sample, ts, context=encoder_hidden_states, fps=None, timestep_cond=timestep_cond)
        return SimpleNamespace(sample=out) if return_dict else (out,)


class DiffusersAutoencoderKL(nn.Module):
    """The slice of diffusers' `AutoencoderKL` the MS pipeline touches (t2v_turbo_ms_pipeline.py:211-217): `.decode(z)[0]`,
    `.config.scaling_factor`, `.dtype`, diffusers-keyed `load_state_dict` — backed by the B200 KL-VAE."""

    def __init__(self, ddconfig=None, scaling_factor=0.18215):
        super().__init__()
        self.vae = AutoencoderKL(ddconfig or VC2_VAE_DDCONFIG, 4)
        self.config = SimpleNamespace(scaling_factor=scaling_factor)

    @property
    def dtype(self):
        return next(self.vae.parameters()).dtype

    def load_state_dict(self, sd, strict=True):
        if any(".mid_block." in k or ".up_blocks." in k for k in sd):
            sd = convert_diffusers_vae_state_dict(sd)
        return self.vae.load_state_dict(sd, strict=strict)

    def decode(self, z, return_dict=False, **kw):
        out = self.vae.decode(z)
        return SimpleNamespace(sample=out) if return_dict else (out,)

    def decode_video(self, z):
        """all frames of [B, 4, F, h, w] in one batched decode (the reference loops over frames)."""
        return self.vae.decode_frames(z, 1.0)


class T2VTurboMSPipeline:
    """pipeline/t2v_turbo_ms_pipeline.py:14-221 (prompt_embeds path; the CLIP text encoder is outside the hot path)."""

    def __init__(self, unet: UNet3DConditionModel, vae: DiffusersAutoencoderKL, text_encoder=None, tokenizer=None, scheduler=None,
                 use_cuda_graph: bool = True):
        self.unet, self.vae, self.text_encoder, self.tokenizer, self.scheduler = unet, vae, text_encoder, tokenizer, scheduler
        self.vae_scale_factor = 8
        self.use_cuda_graph = use_cuda_graph
        self._graph = None

    @property
    def _execution_device(self):
        return next(self.unet.parameters()).device

    def prepare_latents(self, batch_size, num_channels_latents, frames, height, width, dtype, device, generator, latents=None):
        shape = (batch_size, num_channels_latents, frames, height // self.vae_scale_factor, width // self.vae_scale_factor)
        if latents is None:
            gen_dev = generator.device if generator is not None else device
            latents = torch.randn(shape, generator=generator, device=gen_dev, dtype=dtype).to(device)
        else:
            latents = latents.to(device)
        return latents * self.scheduler.init_noise_sigma

    @staticmethod
    def get_w_embedding(w, embedding_dim=512, dtype=torch.float32):
        assert len(w.shape) == 1
        w = w * 1000.0
        half_dim = embedding_dim // 2
        emb = torch.log(torch.tensor(10000.0)) / (half_dim - 1)
        emb = torch.exp(torch.arange(half_dim, dtype=dtype) * -emb)
        emb = w.to(dtype)[:, None] * emb[None, :]
        return torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)

    def _unet_call(self, latents, ts, w_emb, ctx):
        if not self.use_cuda_graph:
            return self.unet(latents, ts, timestep_cond=w_emb, encoder_hidden_states=ctx).sample
        key = (self.unet.model.weight_generation, latents.dtype)
        if self._graph is None or self._graph[0] != key:
            self._graph = (key, _GraphedCall(lambda x, t, w, c: self.unet(x, t, timestep_cond=w, encoder_hidden_states=c).sample))
        return self._graph[1](latents, ts, w_emb, ctx)

    @torch.no_grad()
    def __call__(self, prompt=None, height: Optional[int] = 256, width: Optional[int] = 256, frames: int = 16, guidance_scale: float = 7.5,
                 num_videos_per_prompt: Optional[int] = 1, generator=None, latents=None, num_inference_steps: int = 4,
                 lcm_origin_steps: int = 50, prompt_embeds=None, output_type: Optional[str] = "pil"):
        if prompt_embeds is None:
            raise NotImplementedError("T2VTurboMSPipeline(B200): pass prompt_embeds (the CLIP text encoder is outside the hot path)")
        device = self._execution_device
        prompt_embeds = prompt_embeds.to(device).repeat(1, num_videos_per_prompt, 1).view(-1, prompt_embeds.shape[1], prompt_embeds.shape[2])
        bs = prompt_embeds.shape[0]
        self.scheduler.set_timesteps(num_inference_steps, lcm_origin_steps)
        latents = self.prepare_latents(bs, self.unet.config.in_channels, frames, height, width, prompt_embeds.dtype, device, generator, latents)
        w_embedding = self.get_w_embedding(torch.tensor(guidance_scale).repeat(bs), embedding_dim=256).to(device)
        ctx = prompt_embeds.float()
        denoised = None
        for i, t in enumerate(self.scheduler.timesteps):
            ts = torch.full((bs,), int(t), device=device, dtype=torch.long)
            model_pred = self._unet_call(latents, ts, w_embedding, ctx)
            latents, denoised = self.scheduler.step(model_pred, i, t, latents, generator=generator, return_dict=False)
        if output_type == "latent":
            return denoised
        return self.vae.decode_video(denoised.to(self.vae.dtype) / self.vae.config.scaling_factor)
