"""Host side of the DPM-Solver++(2M) scheduler the reference runners install (generation/lvd.py:46,
diffusers==0.27.2 DPMSolverMultistepScheduler).  Scalar schedule arithmetic lives here (Python floats); the per-element
update runs in the fused CUDA kernel `lvd_cfg_dpm_step` together with the classifier-free-guidance combine.

Interface mirrors what the reference touches: `from_config`, `set_timesteps`, `timesteps`, `scale_model_input`,
`alphas_cumprod` (CPU fp32, indexed by integer timestep; models/pipelines.py:127), `init_noise_sigma`, `order`, `step`.
"""
import math

import numpy as np
import torch

from . import _lib, ops


class DPMSolverMultistepScheduler:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 steps_offset=1, timestep_spacing="leading", prediction_type="epsilon", solver_order=2, **unused):
        if beta_schedule != "scaled_linear" or prediction_type != "epsilon" or solver_order != 2:
            raise ValueError("lvd_b200 scheduler implements the configuration the LVD runners use "
                             "(scaled_linear betas, epsilon prediction, 2nd-order dpmsolver++ midpoint)")
        self.config = dict(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                           beta_schedule=beta_schedule, steps_offset=steps_offset, timestep_spacing=timestep_spacing,
                           prediction_type=prediction_type, solver_order=solver_order)
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.timesteps = None
        self._reset()

    @classmethod
    def from_config(cls, config, **kw):
        cfg = dict(config) if config is not None else {}
        cfg.update(kw)
        keys = ("num_train_timesteps", "beta_start", "beta_end", "beta_schedule", "steps_offset", "timestep_spacing",
                "prediction_type", "solver_order")
        return cls(**{k: cfg[k] for k in keys if k in cfg})

    def _reset(self):
        self._x0_prev = None
        self._lower_order_nums = 0
        self._step_index = 0

    def set_timesteps(self, num_inference_steps, device=None):
        n = int(num_inference_steps)
        nt = self.config["num_train_timesteps"]
        if self.config["timestep_spacing"] == "leading":
            ts = (np.arange(0, n + 1) * (nt // (n + 1))).round()[::-1][:-1].copy().astype(np.int64)
            ts = ts + self.config["steps_offset"]
        elif self.config["timestep_spacing"] == "linspace":
            ts = np.linspace(0, nt - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64)
        else:
            raise ValueError(self.config["timestep_spacing"])
        sig = np.array(((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5)
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        self.sigmas = np.concatenate([sig, [0.0]]).astype(np.float32)
        self.timesteps = torch.from_numpy(ts)
        if device is not None:
            self.timesteps = self.timesteps.to(device)
        self._reset()

    def scale_model_input(self, sample, timestep=None):
        return sample

    @staticmethod
    def _alpha_sigma(sigma):
        alpha = 1.0 / math.sqrt(sigma * sigma + 1.0)
        return alpha, sigma * alpha

    def coefficients(self, i):
        n = len(self.sigmas) - 1
        order = 1 if (self._lower_order_nums < 1 or i == n - 1) else 2
        a_s0, s_s0 = self._alpha_sigma(float(self.sigmas[i]))
        a_t, s_t = self._alpha_sigma(float(self.sigmas[i + 1]))
        lam_s0 = math.log(a_s0) - math.log(s_s0)
        lam_t = math.log(a_t) - math.log(s_t) if s_t > 0 else float("inf")
        h = lam_t - lam_s0
        em1 = math.exp(-h) - 1.0
        c_d1, inv_r0 = 0.0, 0.0
        if order == 2:
            a_s1, s_s1 = self._alpha_sigma(float(self.sigmas[i - 1]))
            inv_r0 = h / (lam_s0 - (math.log(a_s1) - math.log(s_s1)))
            c_d1 = 0.5 * a_t * em1
        return order, s_s0, a_s0, s_t / s_s0, a_t * em1, c_d1, inv_r0

    def step_cfg(self, eps_uncond, eps_cond, sample, guidance_scale):
        """Fused CFG combine + scheduler step on fp16 CUDA tensors of identical shape; returns prev_sample."""
        order, sigma_t, alpha_t, c_x, c_d0, c_d1, inv_r0 = self.coefficients(self._step_index)
        out = torch.empty_like(sample)
        x0 = torch.empty_like(sample)
        assert sample.is_contiguous() and eps_uncond.is_contiguous() and sample.dtype == torch.float16
        _lib.check(_lib.lib().lvd_cfg_dpm_step(
            eps_uncond.data_ptr(), eps_cond.data_ptr() if eps_cond is not None else None, sample.data_ptr(),
            self._x0_prev.data_ptr() if self._x0_prev is not None else None, out.data_ptr(), x0.data_ptr(),
            sample.numel(), float(guidance_scale), sigma_t, alpha_t, c_x, c_d0, c_d1, inv_r0, order, ops._stream()),
            "cfg_dpm_step")
        self._x0_prev = x0
        if self._lower_order_nums < 2:
            self._lower_order_nums += 1
        self._step_index += 1
        return out

    def step(self, model_output, timestep, sample, **kw):
        """diffusers-style signature (controllable_pipeline_text_to_video_synth.py:941-943)."""
        class _Out:
            pass
        o = _Out()
        o.prev_sample = self.step_cfg(model_output.contiguous(), None, sample.contiguous(), 1.0)
        return o
