"""Ito / kappa composition of a shape expert and a colour expert on the probability-flow ODE.

Drop-in for ``sample_composed_ito_ode(shape_model, color_model, shape_label, color_label, args)`` of
``shapes/compose_images_ito.py:88-137`` (``variant="beta"``: divergence of the 1-channel expert w.r.t. its
grayscale input, scaled by 3, update with beta(t)) and of ``shapes/compose_images_ito_2.py:101-151``
(``variant="g2"``: divergence through Grayscale w.r.t. the RGB input, update with g2(t)); ``args`` needs
``.bs .img_size .n_steps``.  The reference gets v^T J v from an autograd VJP through the whole UNet; here the
tangent is pushed forward through the same CUDA kernels (``cdm_unet_forward_jvp``), then ONE fused kernel does
the per-sample kappa reductions and the Euler step.  Also: the 2-D latent variants
(``shapes/visualize_composition_latent_ito.py`` / ``_ito_2.py``) as ``sample_latent_ito_ode``.
"""
import torch

from . import schedule, steps


class Config:
    DEVICE = "cuda"
    SHAPES = ["circle", "square", "triangle"]
    COLORS = ["red", "green", "blue"]


def vector_field(model, t, x, y, probe=None):
    """eps_hat and the Hutchinson divergence estimate v^T J v (``vector_field``, compose_images_ito.py:46-63)."""
    t_in = torch.full((x.shape[0],), t, device=x.device, dtype=torch.float32)
    v = torch.randn_like(x) if probe is None else probe.to(x.device)
    return model.forward_jvp(x, t_in, y, v)


@torch.no_grad()
def sample_composed_ito_ode(shape_model, color_model, shape_label, color_label, args, variant="beta", x_init=None,
                            probes=None):
    device = Config.DEVICE
    shape_model.eval()
    color_model.eval()
    x = (torch.randn(args.bs, 3, args.img_size, args.img_size, device=device) if x_init is None
         else x_init.to(device).float().clone())
    n = args.n_steps
    dt = 1.0 / n
    t_all = torch.tensor([1.0 - i * dt for i in range(n)], dtype=torch.float32)
    sig = schedule.sigma(t_all).tolist()
    a = schedule.dlog_alphadt(t_all).tolist()
    coef = (0.5 * (schedule.beta(t_all) if variant == "beta" else schedule.g2(t_all))).tolist()
    for i in range(n):
        t_val = 1.0 - i * dt
        t = torch.full((x.shape[0],), t_val, device=x.device)
        if variant == "beta":
            x_gray = steps.grayscale(x)
            pv = torch.randn_like(x_gray) if probes is None else probes[i][0].to(device)
            eps_s, div_s = shape_model.forward_jvp(x_gray, t, shape_label, pv)
            scale = 3.0
        else:
            pv3 = torch.randn_like(x) if probes is None else probes[i][0].to(device)
            eps_s, div_s = shape_model.forward_jvp(steps.grayscale(x), t, shape_label, steps.grayscale(pv3),
                                                   pv3.sum(dim=1, keepdim=True))
            scale = 1.0
        pc = torch.randn_like(x) if probes is None else probes[i][1].to(device)
        eps_c, div_c = color_model.forward_jvp(x, t, color_label, pc)
        x = steps.step_ode_kappa(x, eps_s, eps_c, div_s, div_c, sig[i], a[i], coef[i], dt, mode=0, div1_scale=scale, out=x)
    return x


@torch.no_grad()
def sample_latent_ito_ode(model1, model2, n_samples, n_steps, variant="stable", device="cuda", x_init=None, probes=None):
    """2-D latent Ito ODE.  variant "stable": shapes/visualize_composition_latent_ito.py:117-147 (Gaussian probes);
    "clipped": shapes/visualize_composition_latent_ito_2.py:93-119 (jax-faithful schedule, kappa clipped to [-1, 2]).
    model(t, x) are native MLP experts."""
    x = torch.randn(n_samples, 2, device=device) if x_init is None else x_init.to(device).float().clone()
    dt = 1.0 / n_steps
    t_all = torch.tensor([1.0 - i * dt for i in range(n_steps)], dtype=torch.float32)
    a = schedule.dlog_alphadt(t_all).tolist()
    if variant == "stable":
        sig, coef, mode, den = schedule.stable_sigma(t_all).tolist(), schedule.stable_beta(t_all).tolist(), 2, 1e-9
    else:
        jf = schedule.jax_faithful
        sig, coef, mode, den = jf.sigma(t_all).tolist(), jf.beta(t_all).tolist(), 1, 1e-5
    for i in range(n_steps):
        t = torch.full((x.shape[0],), 1.0 - i * dt, device=x.device)
        p1 = torch.randn_like(x) if probes is None else probes[i][0].to(device)
        e1, d1 = model1.forward_jvp(t, x, p1)
        p2 = torch.randn_like(x) if probes is None else probes[i][1].to(device)
        e2, d2 = model2.forward_jvp(t, x, p2)
        x = steps.step_ode_kappa(x, e1, e2, d1, d2, sig[i], a[i], coef[i], dt, mode=mode, den_eps=den, out=x)
    return x
