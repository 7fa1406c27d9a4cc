"""The v2 full fine-tune step of T2V-Turbo on B200 (train_latent_t2v_turbo_v2.py:945-1276, without the reward models): the
teacher's outputs are PRE-COMPUTED (the latent dataset holds z_t, the conditional / unconditional teacher predictions and the
motion-prior score, preprocess_scripts/preprocess_with_motion_prior.py:392-401), every UNet parameter trains, an EMA copy can
serve as the target network.

    t_{n+k} = ddim_timesteps[index],  t_n = max(t_{n+k} - topk, 0)                                                  (:985-990)
    motion_gs = args.motion_gs where (use_motion_guide and index >= (1 - percentage) * N) else 0  -> embedding      (:1019-1039)
    eps_s   = student(z_t, t_{n+k}, c, w_embedding, motion_gs_embedding)          [grad]   full_train.FullUNet      (:1043-1050)
    pred    = c_skip z_t + c_out x0(eps_s)                                                                          (:1051-1060)
    eps_cfg = eps_c + w (eps_c - eps_u),  x0_cfg = x0_c + w (x0_c - x0_u)        from the stored teacher outputs    (:1173-1211)
    eps_cfg -= motion_gs * sqrt(1 - alpha) * score      alpha = alpha_schedule[t_{n+k}] under the same condition    (:1214-1226)
    x_prev  = DDIM step (x0_cfg, eps_cfg) at index                                                                  (:1231-1233)
    target  = c_skip' x_prev + c_out' x0(target_net(x_prev, t_n, c, w_embedding, motion_gs_embedding))   [no grad]  (:1236-1256)
    loss    = huber(pred, target)  (or l2);  backward;  clip_grad_norm_;  AdamW (two lr groups);  EMA               (:1258-1276)

As in `distill.DistillStep`, every affine combination is per-sample scalars times whole latents: folded on the host in fp64 into
coefficients applied by t2v_scale_add_rows.  Note the reference's own arithmetic, kept as is: the motion term enters eps_cfg
only (x0_cfg is formed before it), and its `alpha` is alpha_schedule = sqrt(alphas_cumprod) (:681-682), i.e. the factor is
sqrt(1 - sqrt(alpha_bar)).
"""
from __future__ import annotations

import torch

from . import ops
from .distill import DDIMSolver, guidance_scale_embedding, scalings_for_boundary_conditions


class V2Step:
    """One optimisation step's forward + backward.  student: full_train.FullUNet; target_unet: an inference `UNetModel` whose
    parameters are bound to the student's EMA arena (`student.arena.bind(target_unet, student.arena.target)`), or None — the
    student itself is then the target network (`--use_target_unet` off, :1240)."""

    COEFS = ("k_z", "k_e", "cfg_c", "cfg_u", "x0_z", "x0_e", "mg", "one", "dd_x", "dd_e", "tg_x", "tg_e")

    def __init__(self, student, scheduler, *, target_unet=None, num_ddim_timesteps=200, topk=5, w_min=5.0, w_max=15.0, motion_gs=0.05,
                 percentage=0.5, use_motion_cond=True, loss_type="huber", huber_c=0.001, timestep_scaling_factor=10.0,
                 time_cond_proj_dim=256, fps=16):
        self.student, self.scheduler, self.target_unet = student, scheduler, target_unet
        self.n_ddim, self.topk, self.w_min, self.w_max = num_ddim_timesteps, topk, w_min, w_max
        self.motion_gs, self.percentage, self.use_motion_cond = motion_gs, percentage, use_motion_cond
        self.loss_type, self.huber_c, self.ts_scale = loss_type, huber_c, timestep_scaling_factor
        self.cond_dim, self.fps = time_cond_proj_dim, fps
        ac = scheduler.alphas_cumprod.double().cpu()
        self.alpha, self.sigma = ac.sqrt(), (1 - ac).sqrt()                     # alpha_schedule / sigma_schedule (:681-682)
        self.solver = DDIMSolver(scheduler.alphas_cumprod.cpu().numpy(), ddim_timesteps=num_ddim_timesteps)

    def host_draws(self, index, use_motion_guide, fixed=None, generator=None):
        """index [B] long, use_motion_guide [B] bool (both from the batch) -> dict of CPU tensors: timesteps, w, the two guidance
        embeddings and the COEFS ([B] fp32 each).  `fixed` may pin w [B]."""
        fixed = fixed or {}
        index = index.cpu().long()
        bsz = index.numel()
        start_t = self.solver.ddim_timesteps[index]
        t_n = torch.clamp(start_t - self.topk, min=0)                                                # :986-990
        cs_s, co_s = scalings_for_boundary_conditions(start_t.double(), timestep_scaling=self.ts_scale)
        cs_n, co_n = scalings_for_boundary_conditions(t_n.double(), timestep_scaling=self.ts_scale)
        w = fixed.get("w")
        if w is None:
            w = (self.w_max - self.w_min) * torch.rand((bsz,), generator=generator) + self.w_min      # :1010
        w = w.cpu().double()
        cond = torch.logical_and(use_motion_guide.cpu().bool(), index >= (1 - self.percentage) * self.n_ddim)   # :1025-1031
        mgs = torch.where(cond, torch.full((bsz,), float(self.motion_gs), dtype=torch.float64), torch.zeros(bsz, dtype=torch.float64))
        a_s, s_s, a_n, s_n = self.alpha[start_t], self.sigma[start_t], self.alpha[t_n], self.sigma[t_n]
        a_eff = torch.where(cond, a_s, torch.ones_like(a_s))                                          # :1215-1225
        a_prev = self.solver.ddim_alpha_cumprods_prev[index].double()
        c = dict(k_z=cs_s + co_s / a_s, k_e=-co_s * s_s / a_s,            # c_skip z + c_out (z - sigma eps) / alpha   (:1051-1060)
                 cfg_c=1.0 + w, cfg_u=-w,                                 # eps_c + w (eps_c - eps_u)                  (:1208-1211)
                 x0_z=1.0 / a_s, x0_e=-s_s / a_s,                         # x0 of the CFG eps (linear in eps)          (:1206-1207)
                 mg=-mgs * (1.0 - a_eff).sqrt(), one=torch.ones(bsz, dtype=torch.float64),   # eps -= mgs sqrt(1 - alpha) score
                 dd_x=a_prev.sqrt(), dd_e=(1.0 - a_prev).sqrt(),          # DDIMSolver.ddim_step                       (:1231)
                 tg_x=cs_n + co_n / a_n, tg_e=-co_n * s_n / a_n)          # the target's boundary parametrisation      (:1248-1256)
        out = {k: v.float().contiguous() for k, v in c.items()}
        out.update(index=index, start_timesteps=start_t, timesteps=t_n, w=w, motion_gs=mgs,
                   w_emb=guidance_scale_embedding(w.float(), embedding_dim=self.cond_dim))
        if self.use_motion_cond:                                                                      # :1033-1039
            out["mg_emb"] = guidance_scale_embedding(mgs.float(), embedding_dim=self.cond_dim)
        return out

    def device_step(self, S, z, eps_c, eps_u, score, prompt_embeds):
        """S: the host_draws tensors on the device; z / eps_c / eps_u / score: fp32 [B, 4, T, H, W] contiguous."""
        mg_emb = S.get("mg_emb")
        student = self.student
        eps_s = student(z, S["start_timesteps"], context=prompt_embeds, fps=self.fps, timestep_cond=S["w_emb"], motion_cond=mg_emb)
        model_pred = ops.scale_add_rows(z, S["k_z"], eps_s.float(), S["k_e"])
        saved = student.detach_tapes()
        eps_cfg = ops.scale_add_rows(eps_c, S["cfg_c"], eps_u, S["cfg_u"])
        x0_cfg = ops.scale_add_rows(z, S["x0_z"], eps_cfg, S["x0_e"])
        eps_m = ops.scale_add_rows(eps_cfg, S["one"], score, S["mg"])
        x_prev = ops.scale_add_rows(x0_cfg, S["dd_x"], eps_m, S["dd_e"])
        if self.target_unet is not None:      # the EMA network's inference forward (fused kernels, bf16), no grad
            tu = self.target_unet
            xp = x_prev.to(torch.bfloat16) if tu.dtype == torch.bfloat16 else x_prev
            eps_t = tu(xp, S["timesteps"], context=prompt_embeds, fps=self.fps, timestep_cond=S["w_emb"], motion_cond=mg_emb).float()
        else:                                 # the student itself, gradient-free (still in training mode, as the reference's unet is)
            eps_t = student(x_prev, S["timesteps"], context=prompt_embeds, fps=self.fps, timestep_cond=S["w_emb"], motion_cond=mg_emb).float()
            student.detach_tapes()
        target = ops.scale_add_rows(x_prev, S["tg_x"], eps_t, S["tg_e"])
        student.restore_tapes(saved)
        if self.loss_type == "l2":
            loss, d_pred = ops.mse_loss_grad(model_pred, target)
        else:
            loss, d_pred = ops.huber_loss_grad(model_pred, target, self.huber_c)
        out_extra = {}
        if getattr(self, "reward", None) is not None:     # optional reward branch (:1043-1099): callable(model_pred) -> (loss, d model_pred),
            r_loss, d_r = self.reward(model_pred)         # e.g. functools.partial(vae_train.reward_gradient, vae, reward_fn=..., frame_idx=...)
            d_pred = d_pred + d_r.to(d_pred.dtype)
            out_extra["reward_loss"] = r_loss
        student.backward(ops.scale_add_rows(d_pred, S["k_e"]))
        return dict(loss=loss, model_pred=model_pred, target=target, x_prev=x_prev, **out_extra)

    def __call__(self, batch, *, fixed=None, generator=None):
        """batch: the v2 latent-dataset record (formats.V2_LATENT_KEYS): index, z_t, cond_teacher_out, uncond_teacher_out, score,
        prompt_emb and (optionally) use_motion_guide — latents [B, 4, T, H, W] in any float dtype."""
        dev = batch["z_t"].device
        umg = batch.get("use_motion_guide")
        if umg is None:          # data/mp4_dataset.py:111-114: a sample without the flag uses the motion guidance
            umg = torch.ones(batch["index"].numel(), dtype=torch.bool)
        H = self.host_draws(batch["index"], umg, fixed, generator)
        S = {k: (v.to(dev) if torch.is_tensor(v) and k not in ("index", "w", "motion_gs") else v) for k, v in H.items()}
        f = lambda t: t.to(dev).float().contiguous()     # noqa: E731
        out = self.device_step(S, f(batch["z_t"]), f(batch["cond_teacher_out"]), f(batch["uncond_teacher_out"]), f(batch["score"]),
                               batch["prompt_emb"].to(dev))
        out.update(start_timesteps=H["start_timesteps"], timesteps=H["timesteps"], w=H["w"], motion_gs=H["motion_gs"])
        return out


class GraphedV2Step:
    """V2Step.device_step captured ONCE as a chain of CUDA graphs and replayed per step — the design of distill.GraphedDistillStep
    (verified on B200 for the v1 step): the eager v2 step is tens of thousands of small launches through ctypes, i.e. host bound; the
    capture is cut wherever the backward reports a block of the gradient arena final (`FullUNet.on_grads_final`), so the bucketed
    all-reduce of the 5.65 GB arena keeps its overlap; per step only the host draws and the batch are copied into static buffers.
    Self-target mode only (`--use_target_unet` off, the script's default): the student's bf16 operands are refreshed IN PLACE, so the
    captured kernels keep reading valid buffers, whereas the EMA network's inference weights are re-packed into new buffers after
    every update.  The reward hook (torch autograd) is not capturable either.  Never run on a GPU."""

    BATCH_KEYS = ("z_t", "cond_teacher_out", "uncond_teacher_out", "score")

    def __init__(self, step: V2Step, batch, reducer=None):
        if step.target_unet is not None:
            raise NotImplementedError("GraphedV2Step: the EMA target re-packs its weights out of place; capture the self-target step "
                                      "(target_unet=None) or run the eager V2Step")
        if getattr(step, "reward", None) is not None:
            raise NotImplementedError("GraphedV2Step: the reward branch runs torch autograd and is not captured; use the eager V2Step")
        self.step, self.reducer = step, reducer
        dev = batch["z_t"].device
        self.bufs = {k: batch[k].to(dev).float().contiguous().clone() for k in self.BATCH_KEYS}
        self.prompt = batch["prompt_emb"].to(dev).clone()
        bsz = batch["index"].numel()
        H = step.host_draws(batch["index"], torch.ones(bsz, dtype=torch.bool))
        self.S = {k: v.to(dev) for k, v in H.items() if torch.is_tensor(v) and k not in ("index", "w", "motion_gs")}
        student = step.student
        student.on_grads_final = None
        student.arena.zero_grad()
        self._run()                                   # warm-up: lazy allocations, kernel attributes, the dropout seed
        torch.cuda.synchronize()
        self.segments = []                            # [(graph, arena offset that is final after it)]
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
            self.out = self._run()
            cur[0].capture_end()
            self.segments.append((cur[0], 0))
        torch.cuda.current_stream().wait_stream(stream)
        student.on_grads_final = None

    def _run(self):
        b = self.bufs
        return self.step.device_step(self.S, b["z_t"], b["cond_teacher_out"], b["uncond_teacher_out"], b["score"], self.prompt)

    def __call__(self, batch, *, fixed=None, generator=None):
        umg = batch.get("use_motion_guide")
        if umg is None:
            umg = torch.ones(batch["index"].numel(), dtype=torch.bool)
        H = self.step.host_draws(batch["index"], umg, fixed, generator)
        for k, buf in self.S.items():
            buf.copy_(H[k], non_blocking=True)
        for k in self.BATCH_KEYS:
            self.bufs[k].copy_(batch[k])
        self.prompt.copy_(batch["prompt_emb"])
        for g, offset in self.segments:
            g.replay()
            if self.reducer is not None:
                self.reducer.ready(offset)
        out = dict(self.out)
        out.update(start_timesteps=H["start_timesteps"], timesteps=H["timesteps"], w=H["w"], motion_gs=H["motion_gs"])
        return out


def attach_ema_target(student, target_unet, dtype=torch.bfloat16):
    """Make `target_unet` (a fresh `UNetModel` of the student's configuration, on the student's device) the EMA network of a
    `FullUNet(..., with_target=True)` (train_latent_t2v_turbo_v2.py:733-746): its parameters become views of the student's target arena —
    which starts as a copy of the student's weights and is advanced by `arena.ema_step` —, it runs the fused inference forward in
    `dtype`, and its packed weights are dropped so that the next forward re-packs from the arena."""
    if student.arena.target is None:
        raise RuntimeError("attach_ema_target: build the student with FullUNet(unet, with_target=True)")
    student.arena.bind(target_unet, student.arena.target)
    target_unet.eval()
    target_unet.dtype = dtype
    target_unet.invalidate_packed()
    return target_unet


def train_step_v2(step: V2Step, batch, *, lr, temporal_lr_scale=1.0, ema_decay=0.95, reducer=None, world=1, max_grad_norm=1.0,
                  weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8, accumulate=False, grad_scale=1.0, **kw):
    """zero_grad -> V2Step -> (bucketed NCCL all-reduce of the 5.65 GB gradient arena) -> clip_grad_norm_ + AdamW over the two
    lr groups -> refresh the bf16 operands -> EMA of the target parameters (:1264-1276).
    Gradient accumulation (`accelerator.accumulate(unet)`, :945; --gradient_accumulation_steps): call with accumulate=True for the
    first N - 1 micro-batches — the backward adds into the arena, no exchange, no optimizer step — and normally for the last one with
    grad_scale = 1 / N (accelerate averages the micro-batch losses); the arena is zeroed only at the start of a new accumulation."""
    graphed = isinstance(step, GraphedV2Step)
    student = step.step.student if graphed else step.student
    arena = student.arena
    if not getattr(arena, "_accumulating", False):
        arena.zero_grad()
    arena._accumulating = bool(accumulate)
    if graphed:
        if accumulate:
            raise NotImplementedError("train_step_v2: gradient accumulation with a GraphedV2Step (its replay drives the reducer)")
    elif reducer is not None:
        student.on_grads_final = None if accumulate else reducer.ready      # exchange only once, on the last micro-batch
    out = step(batch, **kw)
    if accumulate:
        return out
    if reducer is not None:
        reducer.finish()
    arena.adamw_step(lr=lr, temporal_lr_scale=temporal_lr_scale, betas=betas, eps=eps, weight_decay=weight_decay, grad_scale=grad_scale / world,
                     max_grad_norm=max_grad_norm)
    student.refresh()
    target_unet = None if graphed else step.target_unet
    if target_unet is not None:
        arena.ema_step(ema_decay)
        target_unet.invalidate_packed()
    return out
