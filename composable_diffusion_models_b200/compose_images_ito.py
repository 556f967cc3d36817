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
import ctypes as C

import torch

from . import _chain, _lib, schedule, steps
from .models import UNet


class Config:
    DEVICE = "cuda"
    SHAPES = ["circle", "square", "triangle"]
    COLORS = ["red", "green", "blue"]


def vector_field(model, t, x, y, probe=None):
    """eps_hat and the Hutchinson divergence estimate v^T J v (``vector_field``, compose_images_ito.py:46-63)."""
    t_in = torch.full((x.shape[0],), t, device=x.device, dtype=torch.float32)
    v = torch.randn_like(x) if probe is None else probe.to(x.device)
    return model.forward_jvp(x, t_in, y, v)


def _ito_tables(n, variant):
    """Per-step scalars of shapes/compose_images_ito.py:102-131 for every i at once (fp32, reference op order)."""
    dt = 1.0 / n
    t_all = torch.tensor([1.0 - i * dt for i in range(n)], dtype=torch.float32)
    sig = schedule.sigma(t_all)
    a = schedule.dlog_alphadt(t_all)
    coef = 0.5 * (schedule.beta(t_all) if variant == "beta" else schedule.g2(t_all))
    return t_all, sig, a, coef


def _ito_chain(models, labels, x, n, variant, probes, seed):
    """The whole Ito / kappa loop through cdm_unet_sample_ito: ONE host call per chunk of steps.  probes: None with a seed ->
    drawn in the library (Philox, no HBM staging); None without a seed -> ``torch.randn`` per step and expert in the
    reference's draw order, staged a chunk of steps at a time; else probes[i][k] (injected)."""
    lib = _lib.lib()
    x = x.contiguous()
    B, S, K = x.shape[0], x.shape[2], len(models)
    if B == 0:
        return x
    t_all, sig, a, coef = _ito_tables(n, variant)
    tab = torch.stack([t_all, sig, a, coef], dim=1)
    prec = _lib.precision_code(models[0].precision)
    harr, hp = _chain.handle_array(models, x.device)
    yp, keep, _ = _chain.label_arrays(labels, B, x.device)
    pch = [1 if (m.in_channels == 1 and variant == "beta") else 3 for m in models]
    in_lib = probes is None and seed is not None
    per_step = sum(pch) * B * S * S * 4
    chunk = n if in_lib else max(1, min(n, (256 << 20) // max(1, per_step)))
    with torch.cuda.device(x.device):
        ws = _chain.workspace(x.device, lib.cdm_unet_sample_ito_workspace_bytes(hp, K, B, S, prec))
        for i0 in range(0, n, chunk):
            m_ = min(chunk, n - i0)
            pp, pkeep, rng = None, [], None
            if in_lib:
                rng = C.byref(_lib.Rng(int(seed), i0 * K))
            else:
                if probes is None:
                    draws = [[torch.randn(B, pch[k], S, S, device=x.device) for k in range(K)] for _ in range(m_)]
                else:
                    draws = [[probes[i][k].to(x.device) for k in range(K)] for i in range(i0, i0 + m_)]
                pp, pkeep = _chain.ptr_array_or_none([torch.stack([d[k] for d in draws]) for k in range(K)])
            ctab, cptr = _chain.host_coef(tab[i0:i0 + m_])
            _lib.check(lib.cdm_unet_sample_ito(hp, K, _lib.ptr(x), yp, 0 if variant == "beta" else 1, pp, rng, cptr, m_, 1.0 / n, B, S,
                                               prec, _lib.ptr(ws), ws.numel(), _lib.stream_of(x)))
            del ctab, pkeep
    del harr, keep
    return x


@torch.no_grad()
def sample_composed_ito_ode_k(models, labels, args, variant="beta", x_init=None, probes=None, seed=None, use_chain=None):
    """K = 2 .. 4 expert Ito superposition on the probability-flow ODE (BASELINE config 4): the K-expert generalisation of
    ``sample_composed_ito_ode`` -- per step every expert's prediction and Hutchinson divergence, the per-sample kappa from the
    (K-1) x (K-1) equal-density-rate system (``cdm_step_ode_kappa_k``; at K = 2 the reference's closed form) and the Euler step.
    models[k]: native ``UNet`` with 1 input channel (a shape-type expert: reads Grayscale(x), divergence x 3 in the "beta"
    variant) or 3 (a colour-type expert).  probes: None -> Gaussian probes (torch RNG in the per-step loop, the library's
    Philox stream in the chain), or probes[i][k] as in the oracle."""
    device = Config.DEVICE
    for m in models:
        m.eval()
    x = (torch.randn(args.bs, 3, args.img_size, args.img_size, device=device) if x_init is None
         else x_init.to(device).float().clone())
    n, K = args.n_steps, len(models)
    if use_chain is None:
        use_chain = _chain.native_all(models, UNet, x) and (K > 2 or models[-1].in_channels == 3)
    if use_chain:
        return _ito_chain(models, labels, x, n, variant, probes, seed)
    t_all, sig, a, coef = _ito_tables(n, variant)
    sig, a, coef = sig.tolist(), a.tolist(), coef.tolist()
    dt = 1.0 / n
    for i in range(n):
        t = torch.full((x.shape[0],), 1.0 - i * dt, device=x.device)
        x_gray = steps.grayscale(x)
        eps, divs, scale = [], [], []
        for k, m in enumerate(models):
            if m.in_channels == 1 and variant == "beta":
                pv = torch.randn_like(x_gray) if probes is None else probes[i][k].to(device)
                e, d = m.forward_jvp(x_gray, t, labels[k], pv)
                scale.append(3.0)
            elif m.in_channels == 1:
                pv3 = torch.randn_like(x) if probes is None else probes[i][k].to(device)
                e, d = m.forward_jvp(x_gray, t, labels[k], steps.grayscale(pv3), pv3.sum(dim=1, keepdim=True))
                scale.append(1.0)
            else:
                pc = torch.randn_like(x) if probes is None else probes[i][k].to(device)
                e, d = m.forward_jvp(x, t, labels[k], pc)
                scale.append(1.0)
            eps.append(e)
            divs.append(d)
        x = steps.step_ode_kappa_k(x, eps, divs, sig[i], a[i], coef[i], dt, div_scale=scale, out=x)
    return x


def bench_step_k4(dev, B, S=64):
    """One timed step of BASELINE config 4 as it is named: Ito superposition of FOUR shapes experts (two shape-type 1-channel
    + two colour-type 3-channel UNets, synthetic weights) -- returns a callable step(i) for bench.py."""
    torch.manual_seed(0)
    models = [UNet(in_channels=c, num_classes=3).to(dev).eval() for c in (1, 1, 3, 3)]
    labels = [torch.full((B,), v, device=dev) for v in (2, 0, 1, 2)]
    st = {"x": torch.randn(B, 3, S, S, device=dev)}
    args = type("A", (), dict(bs=B, img_size=S, n_steps=1000))()
    lib = _lib.lib()
    t_all, sig, a, coef = _ito_tables(1000, "beta")
    tab = torch.stack([t_all, sig, a, coef], dim=1)
    harr, hp = _chain.handle_array(models, st["x"].device)
    yp, keep, _ = _chain.label_arrays(labels, B, st["x"].device)
    prec = _lib.precision_code(models[0].precision)

    def step(i):
        i %= 1000
        ctab, cptr = _chain.host_coef(tab[i:i + 1])
        x = st["x"]
        with torch.cuda.device(x.device):
            ws = _chain.workspace(x.device, lib.cdm_unet_sample_ito_workspace_bytes(hp, 4, B, S, prec))
            _lib.check(lib.cdm_unet_sample_ito(hp, 4, _lib.ptr(x), yp, 0, None, C.byref(_lib.Rng(3, 4 * i)), cptr, 1, 1e-3, B, S, prec,
                                               _lib.ptr(ws), ws.numel(), _lib.stream_of(x)))
    step.keep = (models, labels, harr, keep, args)
    return step


@torch.no_grad()
def sample_composed_ito_ode(shape_model, color_model, shape_label, color_label, args, variant="beta", x_init=None,
                            probes=None, seed=None, use_chain=None):
    device = Config.DEVICE
    shape_model.eval()
    color_model.eval()
    x = (torch.randn(args.bs, 3, args.img_size, args.img_size, device=device) if x_init is None
         else x_init.to(device).float().clone())
    n = args.n_steps
    if use_chain is None:
        use_chain = (_chain.native_all([shape_model, color_model], UNet, x) and shape_model.in_channels == 1
                     and color_model.in_channels == 3)
    if use_chain:
        return _ito_chain([shape_model, color_model], [shape_label, color_label], x, n, variant, probes, seed)
    dt = 1.0 / n
    t_all, sig, a, coef = _ito_tables(n, variant)
    sig, a, coef = sig.tolist(), a.tolist(), coef.tolist()
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
