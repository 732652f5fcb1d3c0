"""TEST INFRASTRUCTURE: restatement of diffusers==0.27.2 `DPMSolverMultistepScheduler` as the reference configures it
(generation/lvd.py:46: from_config of the checkpoint's DDIM config -> beta scaled_linear 0.00085..0.012, 1000 train
steps, epsilon prediction, steps_offset 1, timestep_spacing "leading"; solver_order 2, dpmsolver++, midpoint,
lower_order_final, final_sigmas_type "zero").  diffusers is not vendored in /root/reference and not installable
offline, so this follows the published algorithm; parity for this piece is "unpinned".
Call sites in the reference: controllable_pipeline_text_to_video_synth.py:719-720,912-914,941-943;
models/pipelines.py:76,124-132.
"""
import math

import numpy as np
import torch


class DPMSolverRef:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1,
                 timestep_spacing="leading"):
        self.n_train = num_train_timesteps
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.steps_offset = steps_offset
        self.timestep_spacing = timestep_spacing
        self.model_outputs = [None, None]
        self.lower_order_nums = 0
        self.step_index = None

    def set_timesteps(self, n):
        if self.timestep_spacing == "leading":
            ratio = self.n_train // (n + 1)
            ts = (np.arange(0, n + 1) * ratio).round()[::-1][:-1].copy().astype(np.int64) + self.steps_offset
        else:
            ts = np.linspace(0, self.n_train - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64)
        sig = np.array(((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5)
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        self.sigmas = np.concatenate([sig, [0.0]]).astype(np.float32)
        self.timesteps = ts
        self.model_outputs = [None, None]
        self.lower_order_nums = 0
        self.step_index = None

    def scale_model_input(self, sample, t=None):
        return sample

    @staticmethod
    def _alpha_sigma(sigma):
        alpha = 1.0 / math.sqrt(sigma * sigma + 1.0)
        return alpha, sigma * alpha

    def coefficients(self, i):
        """Scalars of the update from step index i to i+1: (order, sigma_i (as sigma_t of x0), alpha_i, c_x, c_d0,
        c_d1, inv_r0)."""
        n = len(self.timesteps)
        lower_final = i == n - 1
        order = 1 if (self.lower_order_nums < 1 or lower_final) else 2
        a_s0, s_s0 = self._alpha_sigma(float(self.sigmas[i]))
        a_t, s_t = self._alpha_sigma(float(self.sigmas[i + 1]))
        lam_s0 = math.log(a_s0) - math.log(s_s0)
        lam_t = math.log(a_t) - math.log(s_t) if s_t > 0 else float("inf")
        h = lam_t - lam_s0
        em1 = math.exp(-h) - 1.0
        c_x = s_t / s_s0
        c_d0 = a_t * em1
        c_d1, inv_r0 = 0.0, 0.0
        if order == 2:
            a_s1, s_s1 = self._alpha_sigma(float(self.sigmas[i - 1]))
            lam_s1 = math.log(a_s1) - math.log(s_s1)
            r0 = (lam_s0 - lam_s1) / h
            inv_r0 = 1.0 / r0
            c_d1 = 0.5 * a_t * em1
        return order, s_s0, a_s0, c_x, c_d0, c_d1, inv_r0

    def step(self, eps, sample):
        """eps, sample: tensors of the model dtype.  Mirrors step(): x0 in the model dtype, update in fp32, cast back."""
        if self.step_index is None:
            self.step_index = 0
        i = self.step_index
        order, sigma_t, alpha_t, c_x, c_d0, c_d1, inv_r0 = self.coefficients(i)
        dt = eps.dtype
        f = lambda v: torch.tensor(v, dtype=torch.float32, device=eps.device)  # noqa: E731  (0-dim: no promotion)
        x0 = (sample - f(sigma_t) * eps) / f(alpha_t)
        self.model_outputs = [self.model_outputs[1], x0]
        s32 = sample.to(torch.float32)
        out = f(c_x) * s32 - (f(c_d0) * x0)
        if order == 2:
            d1 = f(inv_r0) * (x0 - self.model_outputs[0])
            out = out - (f(c_d1) * d1)
        if self.lower_order_nums < 2:
            self.lower_order_nums += 1
        self.step_index += 1
        return out.to(dt)


def cfg_combine(eps_u, eps_c, guidance_scale):
    """controllable_pipeline_text_to_video_synth.py:926-930."""
    return eps_u + guidance_scale * (eps_c - eps_u)
