"""Reverse-SDE composition of K experts by a weighted SUM of their noise predictions.

Drop-in for ``mnist/compose_scores.py`` (``main(args)`` with ``.model1_path .model2_path .output_file .w1
.w2 .bs .n_steps .xi``) and for the latent loop of ``mnist/visualize_composition_latent.py:63-87``.
Per step the K experts run on libcdm_b200 and ONE fused kernel does combine + Euler-Maruyama update.
"""
import os

import torch

from . import schedule, steps
from .models import UNet, MLP
from .utils import load_checkpoint


def sde_coefficients(n_steps, xi):
    """Per-step scalars of mnist/compose_scores.py:30-43 for every i at once, fp32, reference op order:
    t, a = dlog_alphadt(t), c = beta(t)/sigma(t), g = sqrt(2*xi*beta(t)) * sqrt(dt).  Returns [n_steps, 4]."""
    dt = 1.0 / n_steps
    t = torch.tensor([1.0 - i * dt for i in range(n_steps)], dtype=torch.float32)
    a = schedule.dlog_alphadt(t)
    c = schedule.beta(t) / schedule.sigma(t)
    g = torch.sqrt(2 * xi * schedule.beta(t)) * torch.sqrt(torch.tensor(dt))
    return torch.stack([t, a, c, g], dim=1).contiguous()


@torch.no_grad()
def sample_composed_sde(experts, weights, bs, shape, n_steps, xi=1.0, device="cuda", x_init=None, noise=None,
                        seed=None, call=None):
    """The hot loop of mnist/compose_scores.py:26-46 for K experts.

    experts[k](x, t) -> eps (``call`` overrides the calling convention, e.g. the MLP's (t, x) order).
    noise: None -> ``torch.randn_like`` per step (the reference's RNG order); a [n_steps, B, ...] tensor or a
    callable i -> tensor -> injected; "kernel" -> drawn inside the fused kernel from (seed, i), zero HBM bytes.
    """
    x = torch.randn(bs, *shape, device=device) if x_init is None else x_init.to(device).float().clone()
    coef = sde_coefficients(n_steps, xi).tolist()
    dt = 1.0 / n_steps
    call = call or (lambda m, xx, tt: m(xx, tt))
    for i in range(n_steps):
        tv, a, c, g = coef[i]
        t = torch.full((x.shape[0],), tv, device=x.device)
        eps = [call(m, x, t) for m in experts]
        if isinstance(noise, str) and noise == "kernel":
            x = steps.step_sde(x, eps, weights, a, c, dt, g, rng=(seed or 0, i), out=x)
        else:
            z = torch.randn_like(x) if noise is None else (noise(i) if callable(noise) else noise[i])
            x = steps.step_sde(x, eps, weights, a, c, dt, g, z=z.to(x.device), out=x)
    return x


@torch.no_grad()
def sample_composed_latent_sde(experts, weights, n_samples, n_steps, xi=1.0, device="cuda", x_init=None, noise=None,
                               seed=None, precision="fp32"):
    """mnist/visualize_composition_latent.py:63-87 in ONE persistent launch (cdm_mlp_sample_sde): the whole
    n_steps chain of every sample stays on-chip.  experts: native ``MLP`` modules.  ``precision="fp16"`` runs the two
    256x256 hidden layers on tcgen05 (cdm_mlp_sample_sde_tc; K <= 2 experts of the reference's shape)."""
    import ctypes as C
    from . import _lib
    lib = _lib.lib()
    dev = torch.device(device)
    x = torch.randn(n_samples, experts[0].num_out, device=dev) if x_init is None else x_init.to(dev).float().clone()
    x = x.contiguous()
    coef = sde_coefficients(n_steps, xi).to(dev)
    handles = (C.c_void_p * len(experts))(*[m._native_handle(x.device).value for m in experts])
    z = None
    rng = None
    if isinstance(noise, str) and noise == "kernel":
        rng = C.byref(_lib.Rng(int(seed or 0), 0))
    else:
        z = (torch.stack([torch.randn_like(x) for _ in range(n_steps)]) if noise is None else noise).to(dev).float().contiguous()
    fn = lib.cdm_mlp_sample_sde_tc if _lib.precision_code(precision) == _lib.PREC_F16 else lib.cdm_mlp_sample_sde
    with torch.cuda.device(dev):
        _lib.check(fn(C.cast(handles, C.POINTER(C.c_void_p)), _lib.farray(weights), len(experts),
                                          _lib.ptr(x), _lib.ptr(z), rng, _lib.ptr(coef), n_steps, 1.0 / n_steps,
                                          x.shape[0], _lib.stream_of(x)))
    return x


def main(args, x_init=None, noise=None):
    device = "cuda"
    model1 = UNet().to(device).eval()
    load_checkpoint(model1, None, args.model1_path, device)
    model2 = UNet().to(device).eval()
    load_checkpoint(model2, None, args.model2_path, device)
    x = sample_composed_sde([model1, model2], [args.w1, args.w2], args.bs, (1, 28, 28), args.n_steps, args.xi,
                            device=device, x_init=x_init, noise=noise)
    out = getattr(args, "output_file", None)
    if out:
        os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
        try:
            from torchvision.utils import save_image
            save_image(x.clamp(-1, 1), out, nrow=8, normalize=True, value_range=(-1, 1))
        except Exception:   # image writers are not part of the sampler path
            torch.save(x.cpu(), out + ".pt")
    return x
