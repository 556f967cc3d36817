"""Noise schedules behind the reference's schedule API.

Module-level functions ``log_alpha alpha log_sigma sigma dlog_alphadt beta g2 q_t`` accept a float or a
tensor and return fp32 tensors, like ``mnist/schedule.py:9-62`` / ``shapes/schedule_2.py:50-62``; the
jax-faithful variants of ``shapes/schedule_jax_faithful.py:21-66`` live in ``jax_faithful``.  ``VPSDE``
mirrors ``src/models/compose_grayscale_object_and_color.py:9-18``.

These are O(1) closed forms evaluated on the host once per sampler step; the fused step kernels take
the resulting scalars as arguments.  Everything is evaluated in fp32 in the reference's operation
order on purpose (e.g. ``1 - exp(2 log_alpha)`` loses digits near t = 0 and parity means losing the
same ones).
"""
import torch

beta_0 = 0.1
beta_1 = 20.0


class VPSchedule:
    """Continuous-time variance-preserving schedule with a linear beta(t) ramp."""

    def __init__(self, b0=beta_0, b1=beta_1, sigma_mode="vp"):
        self.b0, self.b1, self.sigma_mode = b0, b1, sigma_mode

    @staticmethod
    def _t(t):
        return torch.as_tensor(t, dtype=torch.float32)

    def log_alpha(self, t):
        t = self._t(t)
        return -0.5 * t * self.b0 - 0.25 * t.pow(2) * (self.b1 - self.b0)

    def alpha(self, t):
        return torch.exp(self.log_alpha(t))

    def log_sigma(self, t):
        t = self._t(t)
        if self.sigma_mode == "jax":
            return torch.log(t + 1e-9)
        return torch.log(1 - torch.exp(2 * self.log_alpha(t)) + 1e-9) / 2

    def sigma(self, t):
        return torch.exp(self.log_sigma(t))

    def dlog_alphadt(self, t):
        t = self._t(t)
        return -0.5 * self.b0 - 0.5 * t * (self.b1 - self.b0)

    def beta(self, t):
        t = self._t(t)
        if self.sigma_mode == "jax":
            return 1 + 0.5 * t * self.b0 + 0.5 * t.pow(2) * (self.b1 - self.b0)
        return -2 * self.dlog_alphadt(t) * self.sigma(t) ** 2

    def g2(self, t):
        if self.sigma_mode == "jax":
            t = self._t(t)
            s = self.sigma(t)
            return 2 * s * 1.0 + 2 * s.pow(2) * self.dlog_alphadt(t)
        return -2 * self.dlog_alphadt(t)

    def q_t(self, x0, t, eps=None):
        if eps is None:
            eps = torch.randn_like(x0)
        shape = (-1,) + (1,) * (x0.dim() - 1)
        a = self.alpha(t).to(x0.device).view(shape)
        s = self.sigma(t).to(x0.device).view(shape)
        return a * x0 + s * eps, eps


_vp = VPSchedule()
jax_faithful = VPSchedule(sigma_mode="jax")

log_alpha = _vp.log_alpha
alpha = _vp.alpha
log_sigma = _vp.log_sigma
sigma = _vp.sigma
dlog_alphadt = _vp.dlog_alphadt
beta = _vp.beta
g2 = _vp.g2
q_t = _vp.q_t


def stable_sigma(t):
    """``stable_sigma`` of shapes/compose_images_ito.py:25-26."""
    t = torch.as_tensor(t, dtype=torch.float32)
    return torch.sqrt(1 - alpha(t) ** 2)


def stable_beta(t):
    """``stable_beta`` of shapes/compose_images_ito.py:33-34."""
    t = torch.as_tensor(t, dtype=torch.float32)
    return -2 * dlog_alphadt(t) * (stable_sigma(t) ** 2)


class VPSDE:
    """Discrete DDPM tables (T steps, linear betas)."""

    def __init__(self, beta_min: float = 0.0001, beta_max: float = 0.02, num_timesteps: int = 1000, device="cpu"):
        self.beta_min, self.beta_max, self.num_timesteps, self.device = beta_min, beta_max, num_timesteps, device
        self.betas = torch.linspace(beta_min, beta_max, num_timesteps, device=device)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, axis=0)
        self.alphas_cumprod_prev = torch.cat([torch.tensor([1.0], device=device), self.alphas_cumprod[:-1]])
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod)
        self.posterior_variance = self.betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)

    def host_tables(self):
        """fp32 CPU copies, used to feed per-step scalars to the fused kernels without device syncs."""
        if not hasattr(self, "_host"):
            self._host = {k: getattr(self, k).detach().float().cpu() for k in
                          ("betas", "alphas", "alphas_cumprod", "alphas_cumprod_prev",
                           "sqrt_one_minus_alphas_cumprod", "posterior_variance")}
        return self._host
