"""`T2VTurboVC2Pipeline` with the reference's call surface (pipeline/t2v_turbo_vc2_pipeline.py:14-220)
driving the B200 UNet, scheduler-step kernel and batched VAE decode.

`LatentVideoModel` stands in for the slice of `LatentDiffusion` the pipeline touches
(pipeline:27-29,144,216; ddpm3d.py:666-679): `.model.diffusion_model`, `.first_stage_model`,
`.cond_stage_model`, `.temporal_length`, `.scale_factor`, `.decode_first_stage_2DAE`.

B200-first additions (all optional, none change results): the UNet forward of a sampling loop is
captured once into a CUDA graph (static shapes; ~1.3k kernel launches replayed per step) and the
decode is a second graph.  The text-context K/V projection of all 16 cross-attention layers is ONE GEMM on 77 rows;
in graph mode it is part of the captured forward (replayed each step), in eager mode it is cached per context tensor.
Graphs are keyed on everything that changes what they would replay — input signature, fps, motion / no motion,
`unet.dtype`, and the weight generation of the UNet / VAE (bumped by `load_state_dict`, `.to()`, `merge_lora`,
`invalidate_packed`) — so a weight update never replays kernels that point at freed packed weights.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Any, Dict, List, Optional, Union

import torch
import torch.nn as nn

from .scheduler import T2VTurboScheduler


class LatentVideoModel(nn.Module):
    def __init__(self, unet, first_stage_model, cond_stage_model=None, temporal_length=16, scale_factor=0.18215):
        super().__init__()
        self.model = nn.Module()
        self.model.diffusion_model = unet
        self.first_stage_model = first_stage_model
        self.cond_stage_model = cond_stage_model if cond_stage_model is not None else nn.Identity()
        self.temporal_length = temporal_length
        self.scale_factor = scale_factor

    @torch.no_grad()
    def decode_first_stage_2DAE(self, z, **kwargs):
        """ddpm3d.py:666-679: z [b, c, t, h, w] -> [b, 3, t, 8h, 8w]; all frames in one batched decode."""
        return self.first_stage_model.decode_frames(z, 1.0 / self.scale_factor)

    @torch.no_grad()
    def encode_first_stage(self, x, noise=None):
        """ddpm3d.py:558-584 (`get_first_stage_encoding` of the posterior sample): video x [b, 3, t, H, W] ->
        scale_factor * z [b, c, t, H/8, W/8]; all frames in one batched encode.  noise: fp32 [b*t, c, h, w]
        (drawn with torch.randn on the CPU like the reference, distributions.py:38-41, when omitted)."""
        return self.first_stage_model.encode_frames(x, noise=noise, scale=self.scale_factor)

    encode_first_stage_2DAE = encode_first_stage      # ddpm3d.py:586-600: same result, frame loop in the reference

    # parameters T2V-Turbo ADDS to the VC2 UNet; absent from `model.ckpt`, filled by the LoRA / unet.pt load (app.py:245-247)
    _T2V_TURBO_EXTRA = ("time_cond_proj.", "motion_cond_proj.", "combine_proj.")

    def load_vc2_checkpoint(self, ckpt, strict_unet=False):
        """VideoCrafter2 `model.ckpt` key space (common_utils.py:399-411): {"state_dict": {...}} or the bare dict with
        `model.diffusion_model.*`, `first_stage_model.*`, `cond_stage_model.*` and the DDPM schedule buffers.  Loads the
        UNet and the KL-VAE; returns the keys that were not consumed (text encoder, schedule buffers, loss weights).
        Like the reference (`load_state_dict(..., strict=False)`, app.py:245-247) the UNet load tolerates the missing
        T2V-Turbo additions (`time_cond_proj`, `motion_cond_proj`, `combine_proj`) — and nothing else: any other
        missing or unexpected key raises.  strict_unet=True demands those too."""
        sd = torch.load(ckpt, map_location="cpu", weights_only=True) if not isinstance(ckpt, dict) else ckpt
        sd = sd.get("state_dict", sd)
        unet_sd = {k[len("model.diffusion_model."):]: v for k, v in sd.items() if k.startswith("model.diffusion_model.")}
        vae_keys = tuple(self.first_stage_model.state_dict().keys())
        vae_sd = {k[len("first_stage_model."):]: v for k, v in sd.items()
                  if k.startswith("first_stage_model.") and k[len("first_stage_model."):] in vae_keys}
        res = self.model.diffusion_model.load_state_dict(unet_sd, strict=strict_unet)
        if not strict_unet:
            bad = [k for k in res.missing_keys if not k.startswith(self._T2V_TURBO_EXTRA)] + list(res.unexpected_keys)
            if bad:
                raise RuntimeError(f"load_vc2_checkpoint: UNet keys missing / unexpected beyond the T2V-Turbo additions: {bad[:8]}"
                                   f"{' ...' if len(bad) > 8 else ''}")
        self.first_stage_model.load_state_dict(vae_sd, strict=True)
        used = {"model.diffusion_model." + k for k in unet_sd} | {"first_stage_model." + k for k in vae_sd}
        return sorted(k for k in sd if k not in used)


_WARMUP_STREAM: dict = {}


def _warmup_stream(device):
    """One reusable side stream per device for graph warm-ups (a fresh stream per capture would make every
    stream-keyed cache downstream grow without bound)."""
    s = _WARMUP_STREAM.get(device)
    if s is None:
        s = _WARMUP_STREAM[device] = torch.cuda.Stream(device=device)
    return s


class _GraphedCall:
    """Capture fn(*static_inputs) once per input signature; replay with inputs copied into static buffers.
    clone_out: hand the caller a copy of the graph's static output (outputs that escape the pipeline must not alias
    the buffer the next replay overwrites)."""

    def __init__(self, fn, clone_out=False):
        self.fn = fn
        self.cache = {}
        self.clone_out = clone_out

    def __call__(self, *tensors):
        key = tuple((tuple(t.shape), t.dtype) for t in tensors)
        ent = self.cache.get(key)
        if ent is None:
            static_in = [t.clone() for t in tensors]
            s = _warmup_stream(tensors[0].device)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):      # warm-up: lazy packing, workspace allocation, kernel attributes
                    self.fn(*static_in)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                static_out = self.fn(*static_in)
            ent = (g, static_in, static_out)
            self.cache[key] = ent
        g, static_in, static_out = ent
        for dst, src in zip(static_in, tensors):
            dst.copy_(src)
        g.replay()
        return static_out.clone() if self.clone_out else static_out


class T2VTurboVC2Pipeline:
    def __init__(self, pretrained_t2v, scheduler: T2VTurboScheduler, model_config: Dict[str, Any] = None,
                 use_cuda_graph: bool = True):
        self.pretrained_t2v = pretrained_t2v
        self.scheduler = scheduler
        self.vae = pretrained_t2v.first_stage_model
        self.unet = pretrained_t2v.model.diffusion_model
        self.text_encoder = pretrained_t2v.cond_stage_model
        self.model_config = model_config
        self.vae_scale_factor = 8
        self.use_cuda_graph = use_cuda_graph
        self._unet_graph = None
        self._vae_graph = None
        self.launch_counter = None

    # reference: DiffusionPipeline properties
    @property
    def _execution_device(self):
        for p in self.unet.parameters():
            return p.device
        return torch.device("cuda")

    @property
    def device(self):
        return self._execution_device

    @property
    def dtype(self):
        return getattr(self.unet, "dtype", torch.float32)

    def to(self, *a, **k):
        self.pretrained_t2v.to(*a, **k)
        self._unet_graph = self._vae_graph = None
        return self

    def _encode_prompt(self, prompt, device, num_videos_per_prompt, prompt_embeds=None):
        if prompt_embeds is None:
            prompt_embeds = self.text_encoder(prompt)
        prompt_embeds = prompt_embeds.to(device=device)
        bs_embed, seq_len, _ = prompt_embeds.shape
        prompt_embeds = prompt_embeds.repeat(1, num_videos_per_prompt, 1)
        return prompt_embeds.view(bs_embed * num_videos_per_prompt, seq_len, -1)

    def prepare_latents(self, batch_size, num_channels_latents, frames, height, width, dtype, device, generator, latents=None):
        shape = (batch_size, num_channels_latents, frames, height // self.vae_scale_factor, width // self.vae_scale_factor)
        if latents is None:
            gen_dev = generator.device if generator is not None else device
            latents = torch.randn(shape, generator=generator, device=gen_dev, dtype=dtype).to(device)
        else:
            latents = latents.to(device)
        return latents * self.scheduler.init_noise_sigma

    def get_w_embedding(self, w, embedding_dim=512, dtype=torch.float32):
        """pipeline:99-120 — host-side, once per call (plain torch on a [bs] vector)."""
        assert len(w.shape) == 1
        w = w * 1000.0
        half_dim = embedding_dim // 2
        emb = torch.log(torch.tensor(10000.0)) / (half_dim - 1)
        emb = torch.exp(torch.arange(half_dim, dtype=dtype) * -emb)
        emb = w.to(dtype)[:, None] * emb[None, :]
        emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)
        if embedding_dim % 2 == 1:
            emb = torch.nn.functional.pad(emb, (0, 1))
        return emb

    def _unet_call(self, latents, ts, ctx, w_emb, motion, fps):
        if not self.use_cuda_graph:
            return self.unet(latents, ts, context=ctx, fps=fps, timestep_cond=w_emb, motion_cond=motion)
        key = (fps, motion is None, self.unet.dtype, getattr(self.unet, "weight_generation", 0))
        if self._unet_graph is None or self._unet_graph[0] != key:
            if motion is None:
                fn = lambda x, t, c, w: self.unet(x, t, context=c, fps=fps, timestep_cond=w)  # noqa: E731
            else:
                fn = lambda x, t, c, w, m: self.unet(x, t, context=c, fps=fps, timestep_cond=w, motion_cond=m)  # noqa: E731
            self._unet_graph = (key, _GraphedCall(fn))   # model_pred is consumed by scheduler.step before the next replay
        args = (latents, ts, ctx, w_emb) if motion is None else (latents, ts, ctx, w_emb, motion)
        return self._unet_graph[1](*args)

    def _decode(self, denoised):
        if not self.use_cuda_graph:
            return self.pretrained_t2v.decode_first_stage_2DAE(denoised)
        key = getattr(self.vae, "weight_generation", 0)
        if self._vae_graph is None or self._vae_graph[0] != key:
            self._vae_graph = (key, _GraphedCall(lambda z: self.pretrained_t2v.decode_first_stage_2DAE(z), clone_out=True))
        return self._vae_graph[1](denoised)

    @torch.no_grad()
    def __call__(self, prompt: Union[str, List[str]] = None, height: Optional[int] = 320, width: Optional[int] = 512,
                 frames: int = 16, fps: int = 16, guidance_scale: float = 7.5, motion_gs: float = 0.1,
                 use_motion_cond: bool = False, percentage: float = 0.3, num_videos_per_prompt: Optional[int] = 1,
                 generator=None, latents: Optional[torch.Tensor] = None, num_inference_steps: int = 4,
                 lcm_origin_steps: int = 50, prompt_embeds: Optional[torch.Tensor] = None,
                 output_type: Optional[str] = "pil"):
        unet_config = self.model_config["params"]["unet_config"]
        frames = self.pretrained_t2v.temporal_length if frames < 0 else frames
        if prompt is not None and isinstance(prompt, str):
            batch_size = 1
        elif prompt is not None and isinstance(prompt, list):
            batch_size = len(prompt)
        else:
            batch_size = prompt_embeds.shape[0]
        device = self._execution_device
        prompt_embeds = self._encode_prompt(prompt, device, num_videos_per_prompt, prompt_embeds=prompt_embeds)
        self.scheduler.set_timesteps(num_inference_steps, lcm_origin_steps)
        timesteps = self.scheduler.timesteps
        num_channels_latents = unet_config["params"]["in_channels"]
        latents = self.prepare_latents(batch_size * num_videos_per_prompt, num_channels_latents, frames, height, width,
                                       prompt_embeds.dtype, device, generator, latents)
        bs = batch_size * num_videos_per_prompt
        ctx = prompt_embeds.to(self.dtype)
        w = torch.tensor(guidance_scale).repeat(bs)
        w_embedding = self.get_w_embedding(w, embedding_dim=256).to(device).to(self.dtype)
        ms_t_threshold = self.scheduler.config.num_train_timesteps * (1 - percentage)
        denoised = None
        for i, t in enumerate(timesteps):
            ts = torch.full((bs,), int(t), device=device, dtype=torch.long)
            motion = None
            if use_motion_cond:
                motion_gs_pt = torch.tensor(motion_gs).repeat(bs)
                if t < ms_t_threshold:
                    motion_gs_pt = torch.zeros_like(motion_gs_pt)
                motion = self.get_w_embedding(motion_gs_pt, embedding_dim=256, dtype=self.dtype).to(device)
            model_pred = self._unet_call(latents, ts, ctx, w_embedding, motion, fps)
            latents, denoised = self.scheduler.step(model_pred, i, t, latents, generator=generator, return_dict=False)
        if not output_type == "latent":
            videos = self._decode(denoised)
        else:
            videos = denoised
        return videos
