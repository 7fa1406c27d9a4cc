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
        self.alpha_cumprods = torch.from_numpy(alpha_cumprods)

    def ddim_step(self, pred_x0, pred_noise, timestep_index):
        a_prev = self.ddim_alpha_cumprods_prev[timestep_index.cpu()].double()
        dev = pred_x0.device
        return ops.scale_add_rows(pred_x0, a_prev.sqrt().float().to(dev), pred_noise, (1.0 - a_prev).sqrt().float().to(dev))


    def reverse_coefs(self, ts):
        """ddim_reverse_step as two per-sample coefficients (fp64 on the host): x_t = ca * x_prev + cb * eps,
        ca = sqrt(a_next / a), cb = sqrt(1 - a_next) - sqrt(1 - a) * ca with a_next = abar[ts], a = abar[max(ts - step_ratio, 0)]."""
        ts = ts.cpu().long()
        prev = (ts - self.step_ratio).clip(min=0)
        a_next, a = self.alpha_cumprods[ts].double(), self.alpha_cumprods[prev].double()
        ca = (a_next / a).sqrt()
        return ca, (1.0 - a_next).sqrt() - (1.0 - a).sqrt() * ca

    def ddim_reverse_step(self, x_prev, pred_noise, ts):
        """One DDIM INVERSION step (ode_solver/ddim_solver.py:89-97; inverse_ddim.py:46-60, motion_prior_sample.py:27-37)."""
        ca, cb = self.reverse_coefs(ts)
        dev = x_prev.device
        return ops.scale_add_rows(x_prev, ca.float().to(dev), pred_noise, cb.float().to(dev))


class DistillStep:
    """One optimisation step's forward + backward (the caller owns the reducer / optimizer calls, see `train_step`).  With a
    data-parallel reducer: `student.on_grads_final = reducer.ready` (eager) or GraphedDistillStep(step, ..., reducer=...)."""

    def __init__(self, student, teacher, scheduler, *, num_ddim_timesteps=50, topk=20, w_min=5.0, w_max=15.0, loss_type="huber",
                 huber_c=0.001, timestep_scaling_factor=10.0, time_cond_proj_dim=256, fps=16):
        self.student, self.teacher, self.scheduler = student, teacher, scheduler
        self.n_ddim, self.topk, self.w_min, self.w_max = num_ddim_timesteps, topk, w_min, w_max
        self.loss_type, self.huber_c, self.ts_scale = loss_type, huber_c, timestep_scaling_factor
        self.cond_dim, self.fps = time_cond_proj_dim, fps
        ac = scheduler.alphas_cumprod.double().cpu()
        self.alpha, self.sigma = ac.sqrt(), (1 - ac).sqrt()                     # :682-683
        self.solver = DDIMSolver(scheduler.alphas_cumprod.cpu().numpy(), ddim_timesteps=num_ddim_timesteps)

    # ---- host side: the random draws of the step and every per-sample coefficient (fp64 on the host, B floats each)
    COEFS = ("an_a", "an_b", "k_z", "k_e", "cfg_c", "cfg_u", "x0_z", "x0_e", "dd_x", "dd_e", "tg_x", "tg_e")

    def host_draws(self, bsz, fixed=None, generator=None):
        """-> dict of CPU tensors: index, start_timesteps, timesteps, w, w_emb [B, cond_dim] and the COEFS ([B] fp32 each).
        `fixed` may pin index [B] / w [B] (the parity test uses the reference's draws)."""
        fixed = fixed or {}
        index = fixed.get("index")
        if index is None:
            index = torch.randint(0, self.n_ddim, (bsz,), generator=generator)                       # :980-982
        index = index.cpu().long()
        start_t = self.solver.ddim_timesteps[index]
        t_n = torch.clamp(start_t - self.topk, min=0)                                                # :983-987
        cs_s, co_s = scalings_for_boundary_conditions(start_t.double(), timestep_scaling=self.ts_scale)
        cs_n, co_n = scalings_for_boundary_conditions(t_n.double(), timestep_scaling=self.ts_scale)
        w = fixed.get("w")
        if w is None:
            w = (self.w_max - self.w_min) * torch.rand((bsz,), generator=generator) + self.w_min      # :1009
        w = w.cpu().double()
        a_s, s_s, a_n, s_n = self.alpha[start_t], self.sigma[start_t], self.alpha[t_n], self.sigma[t_n]
        a_prev = self.solver.ddim_alpha_cumprods_prev[index].double()
        c = dict(an_a=a_s, an_b=s_s,                                      # z = alpha x + sigma noise                  (add_noise)
                 k_z=cs_s + co_s / a_s, k_e=-co_s * s_s / a_s,            # c_skip z + c_out (z - sigma eps) / alpha   (:1030-1039)
                 cfg_c=1.0 + w, cfg_u=-w,                                 # eps_c + w (eps_c - eps_u)                  (:1156-1158)
                 x0_z=1.0 / a_s, x0_e=-s_s / a_s,                         # x0 of the CFG eps (linear: == x0_c + w (x0_c - x0_u))
                 dd_x=a_prev.sqrt(), dd_e=(1.0 - a_prev).sqrt(),          # DDIMSolver.ddim_step                       (:1162)
                 tg_x=cs_n + co_n / a_n, tg_e=-co_n * s_n / a_n)          # the target's boundary parametrisation      (:1172-1181)
        out = {k: v.float().contiguous() for k, v in c.items()}
        out.update(index=index, start_timesteps=start_t, timesteps=t_n, w=w,
                   w_emb=guidance_scale_embedding(w.float(), embedding_dim=self.cond_dim))
        return out

    # ---- device side: no host synchronisation, no host-dependent control flow => CUDA-graph capturable (GraphedDistillStep)
    def device_step(self, S, latents, noise, prompt_embeds, uncond_prompt_embeds):
        """S: the host_draws tensors on the device.  Runs the whole step up to and including the student backward (the
        gradient-arena hooks fire from inside it)."""
        z = ops.scale_add_rows(latents, S["an_a"], noise, S["an_b"])
        eps_s = self.student(z, S["start_timesteps"], context=prompt_embeds, fps=self.fps, timestep_cond=S["w_emb"])
        model_pred = ops.scale_add_rows(z, S["k_z"], eps_s.float(), S["k_e"])
        saved = self.student.detach_tapes()
        # teacher CFG estimate and one DDIM step (no grad; the frozen UNet's inference path)
        zt = z.to(torch.bfloat16) if self.teacher.dtype == torch.bfloat16 else z
        # (the reference's two calls — conditional and unconditional, :1118-1144 — as ONE forward over the concatenated batch:
        # no layer of the UNet mixes samples, and the batched levels run at a better tile fill)
        nb = zt.shape[0]
        eps_cu = self.teacher(torch.cat([zt, zt], 0), torch.cat([S["start_timesteps"], S["start_timesteps"]], 0),
                              context=torch.cat([prompt_embeds, uncond_prompt_embeds], 0), fps=self.fps).float()
        eps_c, eps_u = eps_cu[:nb].contiguous(), eps_cu[nb:].contiguous()
        eps_cfg = ops.scale_add_rows(eps_c, S["cfg_c"], eps_u, S["cfg_u"])
        x0_cfg = ops.scale_add_rows(z, S["x0_z"], eps_cfg, S["x0_e"])
        x_prev = ops.scale_add_rows(x0_cfg, S["dd_x"], eps_cfg, S["dd_e"])
        # target: the student itself on x_prev at t_n, gradient-free (still in training mode, as the reference's unet is)
        eps_t = self.student(x_prev, S["timesteps"], context=prompt_embeds, fps=self.fps, timestep_cond=S["w_emb"])
        self.student.detach_tapes()
        target = ops.scale_add_rows(x_prev, S["tg_x"], eps_t.float(), S["tg_e"])
        self.student.restore_tapes(saved)
        if self.loss_type == "l2":
            loss, d_pred = ops.mse_loss_grad(model_pred, target)
        else:
            loss, d_pred = ops.huber_loss_grad(model_pred, target, self.huber_c)
        out_extra = {}
        if getattr(self, "reward", None) is not None:     # optional reward branch (:1043-1099): callable(model_pred) -> (loss, d model_pred),
            r_loss, d_r = self.reward(model_pred)         # e.g. functools.partial(vae_train.reward_gradient, vae, reward_fn=..., frame_idx=...)
            d_pred = d_pred + d_r.to(d_pred.dtype)
            out_extra["reward_loss"] = r_loss
        self.student.backward(ops.scale_add_rows(d_pred, S["k_e"]))
        return dict(loss=loss, model_pred=model_pred, target=target, x_prev=x_prev, **out_extra)

    def __call__(self, latents, prompt_embeds, uncond_prompt_embeds, *, fixed=None, generator=None):
        """latents [B, 4, T, H, W] fp32 (already scaled by the VAE factor); returns dict(loss, model_pred, target, x_prev, ...).
        `fixed` may pin the random draws (index [B], noise like latents, w [B])."""
        dev = latents.device
        H = self.host_draws(latents.shape[0], fixed, generator)
        noise = (fixed or {}).get("noise")
        if noise is None:
            noise = torch.randn(latents.shape, device=dev, dtype=torch.float32)                        # :1001
        S = {k: (v.to(dev) if torch.is_tensor(v) and k not in ("index", "w") else v) for k, v in H.items()}
        out = self.device_step(S, latents.float().contiguous(), noise.to(dev).float().contiguous(), prompt_embeds, uncond_prompt_embeds)
        out.update(start_timesteps=H["start_timesteps"], timesteps=H["timesteps"], w=H["w"])
        return out


class GraphedDistillStep:
    """DistillStep.device_step captured ONCE as a chain of CUDA graphs and replayed per step: the eager step is ~12 000 small
    launches through ctypes (host bound: 350 ms per step against ~170 ms of GPU work on a B200).  The capture is cut wherever
    the student backward reports a block of the gradient arena final (StudentUNet.on_grads_final), so the data-parallel
    exchange keeps its overlap: after replaying segment k the reducer all-reduces the buckets that segment completed while
    segment k+1 runs.  Per step only the host draws (a few hundred bytes) and the batch are copied into static buffers."""

    def __init__(self, step: DistillStep, latents, prompt_embeds, uncond_prompt_embeds, reducer=None):
        if getattr(step, "reward", None) is not None:
            raise NotImplementedError("GraphedDistillStep: the reward branch runs torch autograd and is not captured; use the eager DistillStep")
        self.step, self.reducer = step, reducer
        dev = latents.device
        self.lat, self.noise = latents.float().clone(), torch.empty_like(latents, dtype=torch.float32)
        self.prompt, self.uncond = prompt_embeds.clone(), uncond_prompt_embeds.clone()
        H = step.host_draws(latents.shape[0])
        self.S = {k: v.to(dev) for k, v in H.items() if torch.is_tensor(v) and k not in ("index", "w")}
        student = step.student
        self.noise.normal_()
        student.on_grads_final = None
        student.arena.zero_grad()
        step.device_step(self.S, self.lat, self.noise, self.prompt, self.uncond)        # warm-up: lazy allocations, kernel attributes
        torch.cuda.synchronize()
        self.segments = []                      # [(graph, arena offset that is final after it)]
        pool = torch.cuda.graph_pool_handle()
        stream = torch.cuda.Stream()
        stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(stream):
            cur = [torch.cuda.CUDAGraph()]
            cur[0].capture_begin(pool=pool, capture_error_mode="thread_local")

            def cut(offset):
                cur[0].capture_end()
                self.segments.append((cur[0], offset))
                cur[0] = torch.cuda.CUDAGraph()
                cur[0].capture_begin(pool=pool, capture_error_mode="thread_local")
            student.on_grads_final = cut
            self.out = step.device_step(self.S, self.lat, self.noise, self.prompt, self.uncond)
            cur[0].capture_end()
            self.segments.append((cur[0], 0))
        torch.cuda.current_stream().wait_stream(stream)
        student.on_grads_final = None

    def __call__(self, latents, prompt_embeds, uncond_prompt_embeds, *, fixed=None, generator=None):
        H = self.step.host_draws(latents.shape[0], fixed, generator)
        for k, buf in self.S.items():
            buf.copy_(H[k], non_blocking=True)
        self.lat.copy_(latents)
        self.prompt.copy_(prompt_embeds)
        self.uncond.copy_(uncond_prompt_embeds)
        if fixed is not None and fixed.get("noise") is not None:
            self.noise.copy_(fixed["noise"])
        else:
            self.noise.normal_()
        for g, offset in self.segments:
            g.replay()
            if self.reducer is not None:
                self.reducer.ready(offset)
        out = dict(self.out)
        out.update(start_timesteps=H["start_timesteps"], timesteps=H["timesteps"], w=H["w"])
        return out


def train_step(step: DistillStep, latents, prompt_embeds, uncond_prompt_embeds, *, lr, reducer=None, world=1, max_grad_norm=1.0,
               weight_decay=1e-2, betas=(0.9, 0.999), eps=1e-8, **kw):
    """zero_grad -> DistillStep -> (bucketed NCCL all-reduce) -> clip_grad_norm_ + fused AdamW -> refresh the bf16 LoRA operands."""
    student = step.step.student if isinstance(step, GraphedDistillStep) else step.student
    arena = student.arena
    arena.zero_grad()
    out = step(latents, prompt_embeds, uncond_prompt_embeds, **kw)     # the reducer's ready() calls fire during the backward
    if reducer is not None:
        reducer.finish()
    arena.adamw_step(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, grad_scale=1.0 / world, max_grad_norm=max_grad_norm)
    student.refresh()
    return out
