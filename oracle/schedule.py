"""Oracle: noise schedules.  Test infrastructure only.

Continuous-time VP schedule, beta_0 = 0.1, beta_1 = 20:
  reference mnist/schedule.py:9-48 (== shapes/schedule.py),
  g2: shapes/schedule_2.py:50-62,
  jax-faithful sigma/beta/g2: shapes/schedule_jax_faithful.py:21-66,
  "stable_*" inline copies: shapes/compose_images_ito.py:15-35.
Discrete DDPM tables: VPSDE, src/models/compose_grayscale_object_and_color.py:9-18;
  get_coeff, src/compositional_diffusion_with_cross_attention.py:212-219.

All functions evaluate in torch float32 in the same operation order as the
reference so that the fp32 rounding of e.g. ``1 - exp(2 log_alpha)`` near t=0
is reproduced, not "improved".
"""
import torch

BETA_0 = 0.1
BETA_1 = 20.0


def _t32(t):
    return torch.as_tensor(t, dtype=torch.float32)


# --- VP schedule (mnist/schedule.py:9-48) ---------------------------------
def log_alpha(t):
    t = _t32(t)
    return -0.5 * t * BETA_0 - 0.25 * t.pow(2) * (BETA_1 - BETA_0)


def alpha(t):
    return torch.exp(log_alpha(t))


def log_sigma(t):
    t = _t32(t)
    return torch.log(1 - torch.exp(2 * log_alpha(t)) + 1e-9) / 2


def sigma(t):
    return torch.exp(log_sigma(t))


def dlog_alphadt(t):
    t = _t32(t)
    return -0.5 * BETA_0 - 0.5 * t * (BETA_1 - BETA_0)


def beta(t):
    t = _t32(t)
    return -2 * dlog_alphadt(t) * sigma(t) ** 2


def g2(t):
    """shapes/schedule_2.py:50-62."""
    return -2 * dlog_alphadt(t)


# --- "stable" inline variants (shapes/compose_images_ito.py:15-35) ---------
def stable_sigma(t):
    t = _t32(t)
    return torch.sqrt(1 - alpha(t) ** 2)


def stable_beta(t):
    t = _t32(t)
    return -2 * dlog_alphadt(t) * (stable_sigma(t) ** 2)


# --- jax-faithful variants (shapes/schedule_jax_faithful.py:21-66) ---------
def jax_sigma(t):
    t = _t32(t)
    return torch.exp(torch.log(t + 1e-9))


def jax_beta(t):
    t = _t32(t)
    return 1 + 0.5 * t * BETA_0 + 0.5 * t.pow(2) * (BETA_1 - BETA_0)


def jax_g2(t):
    t = _t32(t)
    s = jax_sigma(t)
    return 2 * s * 1.0 + 2 * s.pow(2) * dlog_alphadt(t)


def q_t(x0, t, eps):
    """mnist/schedule.py:51-62 (noise always injected here)."""
    a = alpha(t).view(-1, 1, 1, 1)
    s = sigma(t).view(-1, 1, 1, 1)
    return a * x0 + s * eps, eps


# --- discrete tables -------------------------------------------------------
class VPSDETables:
    """src/models/compose_grayscale_object_and_color.py:9-18."""

    def __init__(self, beta_min=0.0001, beta_max=0.02, num_timesteps=1000):
        self.num_timesteps = num_timesteps
        self.betas = torch.linspace(beta_min, beta_max, num_timesteps)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.alphas_cumprod_prev = torch.cat([torch.tensor([1.0]), self.alphas_cumprod[:-1]])
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod)
        self.posterior_variance = self.betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)


def ddpm_alphas_cumprod(timesteps, beta_start=0.0001, beta_end=0.02):
    """get_coeff, src/compositional_diffusion_with_cross_attention.py:212-219."""
    betas = torch.linspace(beta_start, beta_end, timesteps)
    return torch.cumprod(1.0 - betas, dim=0)
