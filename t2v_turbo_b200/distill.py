"""The v1 consistency-distillation step of T2V-Turbo on B200 (train_t2v_turbo_v1_lora.py:943-1196, without the reward models).

    z_{t_{n+k}} = add_noise(latents, noise, t_{n+k})                                   scheduler.add_noise        (:1002-1006)
    eps_s   = student(z, t_{n+k}, c, w_embedding)            [grad]                     train_unet.StudentUNet     (:1023-1029)
    pred    = c_skip z + c_out x0(eps_s)                                                                          (:1030-1039)
    eps_c, eps_u = teacher(z, t_{n+k}, c), teacher(z, t_{n+k}, "")   [no grad]          unet.UNetModel             (:1108-1152)
    x_prev  = DDIM step from the CFG estimate  x0_c + w (x0_c - x0_u), eps_c + w (eps_c - eps_u)                  (:1154-1162)
    target  = c_skip' x_prev + c_out' x0(student(x_prev, t_n, c, w_embedding))   [no grad]                        (:1164-1181)
    loss    = huber(pred, target)  (or l2)                                                                        (:1183-1188)
    backward through the student, gradient all-reduce, clip, AdamW                                                (:1190-1194)

Every affine combination above is per-sample scalars times whole latents: they are folded on the host (fp64) into two
coefficients per tensor and applied by t2v_scale_add_rows; the loss + its gradient is one kernel; the student backward is
`StudentUNet.backward`; the gradient arena goes through `dist.ArenaReducer` + `LoraArena.adamw_step`.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import ops


def scalings_for_boundary_conditions(timestep, sigma_data=0.5, timestep_scaling=10.0):
    """utils/common_utils.py:87-91."""
    s = timestep_scaling * timestep
    return sigma_data ** 2 / (s ** 2 + sigma_data ** 2), s / (s ** 2 + sigma_data ** 2) ** 0.5


def guidance_scale_embedding(w, embedding_dim=512, dtype=torch.float32):
    """utils/common_utils.py:136-163 (same sin / cos layout as the pipeline's get_w_embedding)."""
    w = w * 1000.0
    half = embedding_dim // 2
    emb = math.log(10000.0) / (half - 1)
    emb = torch.exp(torch.arange(half, dtype=dtype) * -emb)
    emb = w.to(dtype)[:, None] * emb[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)
    if embedding_dim % 2 == 1:
        emb = torch.nn.functional.pad(emb, (0, 1))
    return emb


class DDIMSolver:
    """ode_solver/ddim_solver.py:7-87 (use_scale = False, the only mode the training script allows: :691)."""

    def __init__(self, alpha_cumprods, timesteps=1000, ddim_timesteps=50):
        alpha_cumprods = np.asarray(alpha_cumprods)
        self.step_ratio = timesteps // ddim_timesteps
        ts = (np.arange(1, ddim_timesteps + 1) * self.step_ratio).round().astype(np.int64) - 1
        self.ddim_timesteps = torch.from_numpy(ts).long()
        self.ddim_alpha_cumprods = torch.from_numpy(alpha_cumprods[ts])
        self.ddim_alpha_cumprods_prev = torch.from_numpy(np.asarray([alpha_cumprods[0]] + alpha_cumprods[ts[:-1]].tolist()))

    def ddim_step(self, pred_x0, pred_noise, timestep_index):
        a_prev = self.ddim_alpha_cumprods_prev[timestep_index.cpu()].double()
        dev = pred_x0.device
        return ops.scale_add_rows(pred_x0, a_prev.sqrt().float().to(dev), pred_noise, (1.0 - a_prev).sqrt().float().to(dev))


class DistillStep:
    """One optimisation step's forward + backward (the caller owns the reducer / optimizer calls, see `train_step`)."""

    def __init__(self, student, teacher, scheduler, *, num_ddim_timesteps=50, topk=20, w_min=5.0, w_max=15.0, loss_type="huber",
                 huber_c=0.001, timestep_scaling_factor=10.0, time_cond_proj_dim=256, fps=16):
        self.student, self.teacher, self.scheduler = student, teacher, scheduler
        self.n_ddim, self.topk, self.w_min, self.w_max = num_ddim_timesteps, topk, w_min, w_max
        self.loss_type, self.huber_c, self.ts_scale = loss_type, huber_c, timestep_scaling_factor
        self.cond_dim, self.fps = time_cond_proj_dim, fps
        ac = scheduler.alphas_cumprod.double().cpu()
        self.alpha, self.sigma = ac.sqrt(), (1 - ac).sqrt()                     # :682-683
        self.solver = DDIMSolver(scheduler.alphas_cumprod.cpu().numpy(), ddim_timesteps=num_ddim_timesteps)

    def __call__(self, latents, prompt_embeds, uncond_prompt_embeds, *, fixed=None, generator=None):
        """latents [B, 4, T, H, W] (already scaled by the VAE factor); returns dict(loss, model_pred, target, ...).  `fixed` may
        pin the random draws (index [B] long, noise like latents, w [B]) — the parity test uses the reference's draws."""
        fixed = fixed or {}
        dev = latents.device
        bsz = latents.shape[0]
        index = fixed.get("index")
        if index is None:
            index = torch.randint(0, self.n_ddim, (bsz,), generator=generator)
        index = index.cpu().long()
        start_t = self.solver.ddim_timesteps[index]
        t_n = torch.clamp(start_t - self.topk, min=0)                           # :983-987
        cs_s, co_s = scalings_for_boundary_conditions(start_t.double(), timestep_scaling=self.ts_scale)
        cs_n, co_n = scalings_for_boundary_conditions(t_n.double(), timestep_scaling=self.ts_scale)
        noise = fixed.get("noise")
        if noise is None:
            noise = torch.randn(latents.shape, device=dev, dtype=latents.dtype, generator=generator if (generator is not None and generator.device == dev) else None)
        z = self.scheduler.add_noise(latents, noise.to(dev), start_t.to(dev))
        w = fixed.get("w")
        if w is None:
            w = (self.w_max - self.w_min) * torch.rand((bsz,), generator=generator if (generator is not None and generator.device.type == "cpu") else None) + self.w_min
        w = w.cpu().double()
        w_emb = guidance_scale_embedding(w.float(), embedding_dim=self.cond_dim).to(dev)
        f32 = lambda v: v.float().to(dev)

        # ---- online student prediction (kept for the backward)
        eps_s = self.student(z, start_t.to(dev), context=prompt_embeds, fps=self.fps, timestep_cond=w_emb)
        a_s, s_s = self.alpha[start_t], self.sigma[start_t]
        # c_skip z + c_out (z - sigma eps) / alpha  =  (c_skip + c_out / alpha) z  +  (-c_out sigma / alpha) eps
        k_z, k_e = cs_s + co_s / a_s, -co_s * s_s / a_s
        model_pred = ops.scale_add_rows(z.float(), f32(k_z), eps_s.float(), f32(k_e))
        saved = self.student.detach_tapes()

        # ---- teacher CFG estimate and one DDIM step  (no grad; the frozen UNet's inference path)
        zt = z.to(torch.bfloat16) if self.teacher.dtype == torch.bfloat16 else z
        eps_c = self.teacher(zt, start_t.to(dev), context=prompt_embeds, fps=self.fps).float()
        eps_u = self.teacher(zt, start_t.to(dev), context=uncond_prompt_embeds, fps=self.fps).float()
        eps_cfg = ops.scale_add_rows(eps_c, f32(1.0 + w), eps_u, f32(-w))       # eps_c + w (eps_c - eps_u)
        x0_cfg = ops.scale_add_rows(z.float(), f32(1.0 / a_s), eps_cfg, f32(-s_s / a_s))   # linear in eps: == x0_c + w (x0_c - x0_u)
        x_prev = self.solver.ddim_step(x0_cfg, eps_cfg, index)

        # ---- target: the student itself on x_prev at t_n, no grad (still in training mode, as the reference's unet is)
        eps_t = self.student(x_prev, t_n.to(dev), context=prompt_embeds, fps=self.fps, timestep_cond=w_emb)
        self.student.detach_tapes()
        a_n, s_n = self.alpha[t_n], self.sigma[t_n]
        target = ops.scale_add_rows(x_prev, f32(cs_n + co_n / a_n), eps_t.float(), f32(-co_n * s_n / a_n))
        self.student.restore_tapes(saved)

        # ---- loss and its gradient w.r.t. the student's eps prediction
        if self.loss_type == "l2":
            loss, d_pred = ops.mse_loss_grad(model_pred, target)
        else:
            loss, d_pred = ops.huber_loss_grad(model_pred, target, self.huber_c)
        d_eps = ops.scale_add_rows(d_pred, f32(k_e))
        self.student.backward(d_eps)
        return dict(loss=loss, model_pred=model_pred, target=target, x_prev=x_prev, start_timesteps=start_t, timesteps=t_n, w=w)


def train_step(step: DistillStep, latents, prompt_embeds, uncond_prompt_embeds, *, lr, reducer=None, world=1, max_grad_norm=1.0,
               weight_decay=1e-2, betas=(0.9, 0.999), eps=1e-8, **kw):
    """zero_grad -> DistillStep -> (bucketed NCCL all-reduce) -> clip_grad_norm_ + fused AdamW -> refresh the bf16 LoRA operands."""
    arena = step.student.arena
    arena.zero_grad()
    out = step(latents, prompt_embeds, uncond_prompt_embeds, **kw)
    if reducer is not None:
        reducer.ready(0)
        reducer.finish()
    arena.adamw_step(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, grad_scale=1.0 / world, max_grad_norm=max_grad_norm)
    step.student.refresh()
    return out
