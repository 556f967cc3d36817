"""SuperDiff sampling with the linear-solve kappa ("stochastic AND") or softmax kappa (OR), K <= 4 experts.

Drop-in for ``sample_superdiff(model1, model2, label1_idx, label2_idx, mode='OR', T=1.0, l=0.0)`` of
``src/composing_conditional_diffusion_on_shape_and_color_6_1.py:331-430``, generalised from
batch 1 / two experts to ``batch_size`` independent chains and a list of experts (the K-expert system is "K-1 equal
d log q differences + sum(kappa) = 1"; at K = 2 it is the reference's 2x2 system).  The reference builds the system with
``.item()`` round trips and solves it on the host every step; here the inner products, the solve, the DDPM update and the
log-density accumulation are ONE kernel launch per step (``cdm_step_superdiff_solve``).

Experts are called as ``model(img, t_long, label)`` like the reference's ``SimpleUnet``.  ``x_init=``, ``dw=`` ([T, B, ...] unit
normals of the Brownian increments) and ``noise=`` ([T-1, B, ...]) inject the Gaussian draws; omitted, they are drawn with
``torch.randn`` in the reference's order (per step: dW first, then the step noise).
"""
import torch
import torch.nn.functional as F

from . import steps


class Config:
    DEVICE = "cuda"
    IMG_SIZE = 64
    TIMESTEPS = 500


def ddpm_tables(timesteps):
    """reference :99-112 (host copies; the per-step scalars go to the kernel as arguments)."""
    betas = torch.linspace(0.0001, 0.02, timesteps)
    alphas = 1. - betas
    ac = torch.cumprod(alphas, axis=0)
    acp = F.pad(ac[:-1], (1, 0), value=1.0)
    return dict(betas=betas, alphas=alphas, alphas_cumprod=ac, alphas_cumprod_prev=acp,
                sqrt_recip_alphas=torch.sqrt(1.0 / alphas), sqrt_one_minus_alphas_cumprod=torch.sqrt(1. - ac),
                posterior_variance=betas * (1. - acp) / (1. - ac))


def get_forward_process_params(tb, i, timesteps):
    """reference :296-327: finite-difference f_t coefficient and g_t^2 of the OU SDE (same float / double mix)."""
    dt = 1.0 / timesteps
    ac = tb["alphas_cumprod"]
    alpha_t = ac[i].item()
    alpha_t_prev = ac[i - 1].item() if i > 0 else 1.0
    sigma_t_sq = 1 - alpha_t
    sigma_t_sq_prev = 1 - alpha_t_prev
    log_alpha_t = 0.5 * torch.log(ac[i])
    log_alpha_t_prev = 0.5 * torch.log(ac[i - 1]) if i > 0 else 0.0
    d_log_alpha_dt = (log_alpha_t - log_alpha_t_prev) / dt
    log_sigma_t = 0.5 * torch.log(torch.tensor(sigma_t_sq))
    log_sigma_t_prev = 0.5 * torch.log(torch.tensor(sigma_t_sq_prev)) if i > 0 else torch.tensor(-float('inf'))
    d_log_sigma_dt = (log_sigma_t - log_sigma_t_prev) / dt if torch.isfinite(log_sigma_t_prev) else 0.0
    g_t_sq = 2 * sigma_t_sq * (d_log_sigma_dt - d_log_alpha_dt)
    g_t_sq = max(g_t_sq, 1e-8)
    return float(d_log_alpha_dt), float(g_t_sq)


@torch.no_grad()
def sample_superdiff(model1, model2, label1_idx, label2_idx, mode='OR', T=1.0, l=0.0, models=None, labels=None,   # noqa: E741
                     batch_size=1, x_init=None, dw=None, noise=None, config=Config, return_log_q=False):
    if mode not in ('OR', 'AND'):
        raise ValueError("Mode must be 'OR' or 'AND'")
    dev = config.DEVICE
    ms = list(models) if models is not None else [model1, model2]
    lab_idx = list(labels) if labels is not None else [label1_idx, label2_idx]
    n_t = config.TIMESTEPS
    img = (torch.randn((batch_size, 3, config.IMG_SIZE, config.IMG_SIZE), device=dev) if x_init is None
           else x_init.to(dev).float().clone())
    bs = img.shape[0]
    log_q = torch.zeros(bs, len(ms), device=dev)
    labs = [torch.full((bs,), int(v), device=dev, dtype=torch.long) for v in lab_idx]
    tb = ddpm_tables(n_t)
    d_tau = 1.0 / n_t
    for n, i in enumerate(reversed(range(n_t))):
        t = torch.full((bs,), i, device=dev, dtype=torch.long)
        preds = [m(img, t, lab) for m, lab in zip(ms, labs)]
        f_coef, g_sq = get_forward_process_params(tb, i, n_t)
        dwn = None
        if mode == 'AND':
            dwn = torch.randn_like(img) if dw is None else dw[n].to(dev)
        z = None
        if i > 0:
            z = torch.randn_like(img) if noise is None else noise[n].to(dev)
        img = steps.step_superdiff_solve(img, preds, log_q, mode, T, l, float(tb["sqrt_one_minus_alphas_cumprod"][i]),
                                         float(tb["betas"][i]), float(tb["sqrt_recip_alphas"][i]),
                                         float(torch.sqrt(tb["posterior_variance"][i])) if i > 0 else 0.0, d_tau, f_coef, g_sq,
                                         dw=dwn, z=z, out=img)
    return (img, log_q) if return_log_q else img
