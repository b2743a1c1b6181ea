"""Noise schedule of the variational diffusion process (host side, tiny).

Restates reference src/models/components/variational_diffusion.py:68-107,206-255 (polynomial schedule, gamma
lookup table) and :316-367,1219-1253 (per-step coefficients) so the sampler does not depend on the reference.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def gamma_table(num_timesteps: int = 1000, precision: float = 1e-5, schedule: str = "polynomial_2") -> torch.Tensor:
    """gamma[0..T] (float32), built in float64 numpy exactly like PredefinedNoiseSchedule."""
    if not schedule.startswith("polynomial"):
        raise NotImplementedError(f"noise schedule '{schedule}' (shipped configs use polynomial_2)")
    power = float(schedule.split("_")[1])
    steps = num_timesteps + 1
    x = np.linspace(0, steps, steps)
    alphas2 = (1 - np.power(x / steps, power)) ** 2
    alphas2 = np.concatenate([np.ones(1), alphas2], axis=0)
    step = np.clip(alphas2[1:] / alphas2[:-1], a_min=0.001, a_max=1.0)
    alphas2 = np.cumprod(step, axis=0)
    alphas2 = (1 - 2 * precision) * alphas2 + precision
    sigmas2 = 1 - alphas2
    return torch.tensor(-(np.log(alphas2) - np.log(sigmas2))).float()


def step_coefficient_table(gamma: torch.Tensor, sample_steps: int) -> torch.Tensor:
    """Rows {alpha_ts, c_eps, sigma, t} for s = sample_steps-1 .. 0 (row r is the r-th reverse step).

    t = (s+1)/sample_steps, s/sample_steps index gamma via round(t * T) (variational_diffusion.py:252-255,
    1336-1339); all arithmetic in float32 like the reference.
    """
    T = gamma.shape[0] - 1
    s_int = torch.arange(sample_steps - 1, -1, -1, dtype=torch.float32)
    s = s_int / sample_steps
    t = (s_int + 1) / sample_steps
    g_s = gamma[torch.round(s * T).long()]
    g_t = gamma[torch.round(t * T).long()]
    sigma2_ts = -torch.expm1(F.softplus(g_s) - F.softplus(g_t))
    alpha_ts = torch.exp(0.5 * (F.logsigmoid(-g_t) - F.logsigmoid(-g_s)))
    sigma_ts = torch.sqrt(sigma2_ts)
    sigma_s = torch.sqrt(torch.sigmoid(g_s))
    sigma_t = torch.sqrt(torch.sigmoid(g_t))
    c_eps = sigma2_ts / alpha_ts / sigma_t
    sigma = sigma_ts * sigma_s / sigma_t
    return torch.stack((alpha_ts, c_eps, sigma, t), dim=1).contiguous()


def decode_coefficients(gamma: torch.Tensor) -> torch.Tensor:
    """{1/alpha_0, sigma_0, sigma_x, 0} for p(x,h | z_0) (variational_diffusion.py:559-577, 860-885)."""
    g0 = gamma[0]
    sigma_x = torch.exp(0.5 * g0)
    sigma_0 = torch.sqrt(torch.sigmoid(g0))
    alpha_0 = torch.sqrt(torch.sigmoid(-g0))
    return torch.stack((1.0 / alpha_0, sigma_0, sigma_x, torch.zeros(()))).contiguous()
