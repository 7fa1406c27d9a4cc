"""ORACLE (test infrastructure): restatement of `T2VTurboScheduler` (scheduler/t2v_turbo_scheduler.py)
— scaled-linear betas (:183-233), LCM timestep grid (:323-355), boundary scalings (:359-365) and the
epsilon-parameterised multistep `step` (:367-467) — without the diffusers base classes.
Known answers (produced by executing the reference) are checked in tests/."""
from __future__ import annotations

import numpy as np
import torch


class SchedulerOracle:
    def __init__(self, num_train_timesteps=1000, linear_start=0.00085, linear_end=0.012):
        self.num_train_timesteps = num_train_timesteps
        self.betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0)
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps, lcm_origin_steps):
        c = self.num_train_timesteps // lcm_origin_steps
        grid = np.asarray(list(range(1, lcm_origin_steps + 1))) * c - 1
        skipping = len(grid) // num_inference_steps
        self.timesteps = torch.from_numpy(grid[::-skipping][:num_inference_steps].copy())
        return self.timesteps

    @staticmethod
    def scalings(t, sigma_data=0.5):
        c_skip = sigma_data ** 2 / ((t / 0.1) ** 2 + sigma_data ** 2)
        c_out = (t / 0.1) / ((t / 0.1) ** 2 + sigma_data ** 2) ** 0.5
        return c_skip, c_out

    def step(self, model_output, timeindex, timestep, sample, noise=None):
        prev_idx = timeindex + 1
        prev_t = self.timesteps[prev_idx] if prev_idx < len(self.timesteps) else timestep
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        c_skip, c_out = self.scalings(timestep)
        x0 = (sample - (1 - a_t).sqrt() * model_output) / a_t.sqrt()
        den = c_out * x0 + c_skip * sample
        if len(self.timesteps) > 1:
            prev = a_prev.sqrt() * den + (1 - a_prev).sqrt() * noise
        else:
            prev = den
        return prev, den

    def add_noise(self, original_samples, noise, timesteps):
        """scheduler/t2v_turbo_scheduler.py:470-495: the alpha table is cast to the sample dtype before the square roots."""
        ac = self.alphas_cumprod.to(dtype=original_samples.dtype)[timesteps]
        shape = (-1,) + (1,) * (original_samples.dim() - 1)
        return (ac ** 0.5).view(shape) * original_samples + ((1 - ac) ** 0.5).view(shape) * noise
