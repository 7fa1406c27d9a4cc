"""`T2VTurboScheduler` with the reference's public surface (scheduler/t2v_turbo_scheduler.py:129-495),
without the diffusers base classes (diffusers is not a dependency here).

Host-side schedule tables are plain torch/numpy exactly as in the reference (:183-250, :323-355, :359-365);
the per-step tensor arithmetic (`step`, :367-467) is ONE fused kernel (t2v_lcm_step) when the sample
lives on a CUDA device.  Noise is still drawn with torch.randn from the caller's generator so the
random stream is bit-identical to the reference's `randn_tensor`.
"""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

from . import ops


@dataclass
class T2VTurboSchedulerOutput:
    prev_sample: torch.Tensor
    denoised: Optional[torch.Tensor] = None


class T2VTurboScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, linear_start: float = 0.00085, linear_end: float = 0.012,
                 beta_schedule: str = "scaled_linear", trained_betas=None, clip_sample: bool = True,
                 set_alpha_to_one: bool = True, steps_offset: int = 0, prediction_type: str = "epsilon",
                 thresholding: bool = False, dynamic_thresholding_ratio: float = 0.995, clip_sample_range: float = 1.0,
                 sample_max_value: float = 1.0, timestep_spacing: str = "leading", rescale_betas_zero_snr: bool = False):
        assert beta_schedule == "scaled_linear" and trained_betas is None   # reference :201-202
        if prediction_type != "epsilon" or rescale_betas_zero_snr:
            raise NotImplementedError("T2VTurboScheduler(B200): only the epsilon parameterisation used by T2V-Turbo")
        self.config = SimpleNamespace(
            num_train_timesteps=num_train_timesteps, linear_start=linear_start, linear_end=linear_end,
            beta_schedule=beta_schedule, trained_betas=trained_betas, clip_sample=clip_sample,
            set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset, prediction_type=prediction_type,
            thresholding=thresholding, dynamic_thresholding_ratio=dynamic_thresholding_ratio,
            clip_sample_range=clip_sample_range, sample_max_value=sample_max_value,
            timestep_spacing=timestep_spacing, rescale_betas_zero_snr=rescale_betas_zero_snr)
        self.betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps: int, lcm_origin_steps: int, device=None):
        if num_inference_steps > self.config.num_train_timesteps:
            raise ValueError(f"`num_inference_steps`: {num_inference_steps} cannot be larger than "
                             f"`self.config.train_timesteps`: {self.config.num_train_timesteps}")
        self.num_inference_steps = num_inference_steps
        c = self.config.num_train_timesteps // lcm_origin_steps
        lcm_origin_timesteps = np.asarray(list(range(1, lcm_origin_steps + 1))) * c - 1
        skipping_step = len(lcm_origin_timesteps) // num_inference_steps
        timesteps = lcm_origin_timesteps[::-skipping_step][:num_inference_steps]
        # kept on the host: the reference indexes CPU tables with these values every step
        self.timesteps = torch.from_numpy(timesteps.copy())

    def get_scalings_for_boundary_condition_discrete(self, t):
        self.sigma_data = 0.5
        c_skip = self.sigma_data ** 2 / ((t / 0.1) ** 2 + self.sigma_data ** 2)
        c_out = (t / 0.1) / ((t / 0.1) ** 2 + self.sigma_data ** 2) ** 0.5
        return c_skip, c_out

    def step_coefficients(self, timeindex: int, timestep):
        """The six fp32 scalars of one step (reference :423-447)."""
        timestep = int(timestep)
        prev_timeindex = timeindex + 1
        prev_timestep = int(self.timesteps[prev_timeindex]) if prev_timeindex < len(self.timesteps) else timestep
        alpha_prod_t = self.alphas_cumprod[timestep]
        alpha_prod_t_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        c_skip, c_out = self.get_scalings_for_boundary_condition_discrete(torch.tensor(timestep))
        one = torch.tensor(1.0)
        return dict(inv_sqrt_alpha_t=float(one / alpha_prod_t.sqrt()), sqrt_beta_t=float((1 - alpha_prod_t).sqrt()),
                    c_skip=float(c_skip.float()), c_out=float(c_out.float()),
                    sqrt_alpha_prev=float(alpha_prod_t_prev.sqrt()), sqrt_beta_prev=float((1 - alpha_prod_t_prev).sqrt()))

    def step(self, model_output, timeindex: int, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False,
             generator=None, variance_noise=None, return_dict: bool = True):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if not sample.is_cuda:
            raise RuntimeError("T2VTurboScheduler(B200).step: sample must be a CUDA tensor (no CPU fallback)")
        coef = self.step_coefficients(timeindex, timestep)
        noise = None
        if len(self.timesteps) > 1:
            if variance_noise is not None:
                noise = variance_noise
            else:
                gen_dev = generator.device if generator is not None else sample.device
                noise = torch.randn(sample.shape, generator=generator, device=gen_dev, dtype=sample.dtype).to(sample.device)
        prev_sample, denoised = ops.lcm_step(sample.contiguous(), model_output.contiguous(),
                                             noise.contiguous() if noise is not None else None, **coef)
        if not return_dict:
            return (prev_sample, denoised)
        return T2VTurboSchedulerOutput(prev_sample=prev_sample, denoised=denoised)

    def add_noise(self, original_samples, noise, timesteps):
        """Reference :470-495.  The per-sample coefficients are computed exactly as the reference does (the alpha table
        cast to the sample dtype, then `** 0.5` in that dtype: a [b]-element gather on the device); the tensor
        arithmetic `sqrt_a * x0 + sqrt(1 - a) * noise` is one kernel (t2v_scale_add_rows)."""
        if not original_samples.is_cuda:
            raise RuntimeError("T2VTurboScheduler(B200).add_noise: samples must be CUDA tensors (no CPU fallback)")
        alphas_cumprod = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        sqrt_alpha_prod = (alphas_cumprod[timesteps] ** 0.5).flatten().float().contiguous()
        sqrt_one_minus = ((1 - alphas_cumprod[timesteps]) ** 0.5).flatten().float().contiguous()
        return ops.scale_add_rows(original_samples.contiguous(), sqrt_alpha_prod,
                                  noise.to(original_samples.dtype).contiguous(), sqrt_one_minus)

    def __len__(self):
        return self.config.num_train_timesteps
