"""``DiffusionSDE`` (DDPM schedule + the finite-difference SDE coefficients of the Ito density estimator) and the batched
SuperDiff sampler built on it.

Drop-in for ``src/composing_conditional_diffusion_on_shape_and_color_3.py``: ``DiffusionSDE(timesteps, img_dims, device)``
with the reference's attributes (``betas alphas alphas_cumprod alphas_cumprod_prev sqrt_alphas_cumprod
sqrt_one_minus_alphas_cumprod posterior_variance f_t_coeff g_t_sq div_f_t``, :125-159), ``_extract`` / ``q_sample`` /
``p_sample`` (:161-180), and ``sample_superdiff(shape_model, color_model, diffusion, shape_class_idx, color_class_idx,
num_images=1, strategy='OR', temp=1.0, bias=0.0)`` (:346-430).  Per step the two experts run, then ONE fused kernel
(``cdm_step_superdiff_solve``) does the kappa softmax, the kappa-weighted DDPM update and both log-density increments; the
reference does ~30 tensor ops and two ``_extract`` gathers per quantity.

Note: as shipped the reference function raises at its first log-density update (``div_f_t_t`` is unsqueezed twice, :405);
this implements what the surrounding arithmetic evidently means, pinned by ``tests/golden/sampler_superdiff3_*.npz``.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import steps


class Config:
    DEVICE = "cuda"
    IMG_SIZE = 64
    TIMESTEPS = 500


def get_linear_beta_schedule(timesteps):
    return torch.linspace(0.0001, 0.02, timesteps)


class DiffusionSDE:
    def __init__(self, timesteps, img_dims, device):
        self.timesteps, self.device, self.img_dims = timesteps, device, img_dims
        host = {}
        host["betas"] = get_linear_beta_schedule(timesteps)
        host["alphas"] = 1. - host["betas"]
        ac = host["alphas_cumprod"] = torch.cumprod(host["alphas"], axis=0)
        host["alphas_cumprod_prev"] = F.pad(ac[:-1], (1, 0), value=1.0)
        host["sqrt_alphas_cumprod"] = torch.sqrt(ac)
        host["sqrt_one_minus_alphas_cumprod"] = torch.sqrt(1. - ac)
        host["posterior_variance"] = host["betas"] * (1. - host["alphas_cumprod_prev"]) / (1. - ac)
        # OU SDE dx = f_t x dt + g_t dW: backward differences of log alpha_t / log sigma_t, zero-padded at t = 0, times T
        log_alpha_t = 0.5 * torch.log(ac)
        log_sigma_t = 0.5 * torch.log(1. - ac)
        host["f_t_coeff"] = (log_alpha_t - F.pad(log_alpha_t[:-1], (1, 0))) * timesteps
        d_ls = ((log_sigma_t - log_alpha_t) - F.pad((log_sigma_t - log_alpha_t)[:-1], (1, 0))) * timesteps
        host["g_t_sq"] = 2 * (1. - ac) * d_ls
        host["div_f_t"] = np.prod(img_dims) * host["f_t_coeff"]
        self._host = host                       # host copies: the per-step scalars go to the kernels as arguments
        for k, v in host.items():
            setattr(self, k, v.to(device))

    def host_tables(self):
        return self._host

    def _extract(self, a, t, x_shape):
        out = a.gather(-1, t)
        return out.reshape(t.shape[0], *((1,) * (len(x_shape) - 1)))

    def q_sample(self, x_0, t, noise=None):
        if noise is None:
            noise = torch.randn_like(x_0)
        return (self._extract(self.sqrt_alphas_cumprod, t, x_0.shape) * x_0
                + self._extract(self.sqrt_one_minus_alphas_cumprod, t, x_0.shape) * noise)

    @torch.no_grad()
    def p_sample(self, model_output, x, t, noise=None):
        """One DDPM ancestral step (:167-180) as a single fused launch; every sample shares t (as in the reference's loops)."""
        i = int(t[0])
        h = self._host
        z = None
        if i > 0:
            z = torch.randn_like(x) if noise is None else noise
        return steps.step_cfg(x, [model_output], [1.0], 1.0, 1, 1, float(torch.sqrt(1.0 / h["alphas"])[i]), float(h["betas"][i]),
                              float(h["sqrt_one_minus_alphas_cumprod"][i]), float(torch.sqrt(h["posterior_variance"][i])), z=z)


@torch.no_grad()
def sample_superdiff(shape_model, color_model, diffusion, shape_class_idx, color_class_idx, num_images=1, strategy='OR',
                     temp=1.0, bias=0.0, x_init=None, noise=None, return_log_q=False):
    device = diffusion.device
    x = (torch.randn((num_images, 3, Config.IMG_SIZE, Config.IMG_SIZE), device=device) if x_init is None
         else x_init.to(device).float().clone())
    bs = x.shape[0]
    c_shape = torch.full((bs,), int(shape_class_idx), device=device, dtype=torch.long)
    c_color = torch.full((bs,), int(color_class_idx), device=device, dtype=torch.long)
    log_q = torch.zeros(bs, 2, device=device)
    h = diffusion.host_tables()
    T = diffusion.timesteps
    sra = torch.sqrt(1.0 / h["alphas"])
    spv = torch.sqrt(h["posterior_variance"])
    # strategy 'OR': kappa = softmax(temp * log_q + bias); anything else: 0.5 / 0.5 == the softmax of zeros
    tk, bk = (temp, bias) if strategy == 'OR' else (0.0, 0.0)
    for n, i in enumerate(range(T - 1, -1, -1)):
        t = torch.full((bs,), i, device=device, dtype=torch.long)
        eps = [shape_model(x, t, c_shape), color_model(x, t, c_color)]
        z = None
        if i > 0:
            z = torch.randn_like(x) if noise is None else noise[n].to(device)
        # scores are -eps / (sigma_t + 1e-8) (:401-402); the kernel divides by one denominator throughout, which moves the
        # update's beta * eps / sigma_t term by <= 1e-8 / sigma_t relative
        som = float(h["sqrt_one_minus_alphas_cumprod"][i] + 1e-8)
        x = steps.step_superdiff_solve(x, eps, log_q, "OR", tk, bk, som, float(h["betas"][i]), float(sra[i]),
                                       float(spv[i]) if i > 0 else 0.0, 1.0 / T, float(h["f_t_coeff"][i]), float(h["g_t_sq"][i]),
                                       z=z, out=x)
    return (x, log_q) if return_log_q else x
