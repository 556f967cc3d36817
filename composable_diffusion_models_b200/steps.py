"""Thin tensor-level wrappers over the fused step kernels of libcdm_b200 (rows a9-a13).

Each function enqueues ONE kernel on the current CUDA stream.  ``x`` is [B, C, *spatial] (or [B, D] for
latents) fp32 on a CUDA device; expert outputs in ``eps`` have either C channels or 1 (broadcast).
``z`` is an injected noise tensor, or ``rng=(seed, step)`` asks the kernel to draw its own.
"""
import ctypes as C

import torch

from . import _lib


def _shape3(x):
    if x.dim() == 2:
        return x.shape[0], 1, x.shape[1]
    B, Cc = x.shape[0], x.shape[1]
    hw = 1
    for s in x.shape[2:]:
        hw *= s
    return B, Cc, hw


def _channels(eps, x):
    if x.dim() == 2:
        return [1] * len(eps)
    out = []
    for e in eps:
        if e.shape[0] != x.shape[0] or e.shape[2:] != x.shape[2:] or e.shape[1] not in (1, x.shape[1]):
            raise ValueError(f"expert output {tuple(e.shape)} does not match x {tuple(x.shape)}")
        out.append(e.shape[1])
    return out


def _prep(x, eps, z):
    _lib.require_cuda(x, z, *eps)
    x = x.contiguous()
    eps = [e.float().contiguous() for e in eps]
    z = z.float().contiguous() if z is not None else None
    return x, eps, z


def _call(ref, name, *args):
    """Run one C-ABI entry with ``ref``'s device current: the library launches on the CURRENT device with the stream
    handle it is given, so a tensor on cuda:1 must not be stepped while cuda:0 is current."""
    with torch.cuda.device(ref.device):
        _lib.check(getattr(_lib.lib(), name)(*args))


def _rng(rng):
    return C.byref(_lib.Rng(int(rng[0]), int(rng[1]))) if rng is not None else None


def step_sde(x, eps, weights, a, c, dt, g, z=None, rng=None, out=None):
    """mnist/compose_scores.py:37-46:  x + (-(a*x - c*sum_k w_k eps_k)*dt + g*z)."""
    x, eps, z = _prep(x, eps, z)
    B, Cc, HW = _shape3(x)
    out = torch.empty_like(x) if out is None else out
    _call(x, "cdm_step_sde", _lib.ptr(x), _lib.ptr_array(eps), _lib.iarray(_channels(eps, x)),
                                       _lib.farray(weights), len(eps), _lib.ptr(z), _rng(rng), a, c, dt, g,
                                       _lib.ptr(out), B, Cc, HW, _lib.stream_of(x))
    return out


def step_ddim(x, eps, weights, wsum, alpha_now, sigma_now, alpha_next, sigma_next, out=None, gray_out=None):
    """shapes/compose_images_ddim.py:52-68 (+ Grayscale of the result for the next step, :47)."""
    x, eps, _ = _prep(x, eps, None)
    B, Cc, HW = _shape3(x)
    out = torch.empty_like(x) if out is None else out
    _call(x, "cdm_step_ddim", _lib.ptr(x), _lib.ptr_array(eps), _lib.iarray(_channels(eps, x)),
                                        _lib.farray(weights), len(eps), wsum, alpha_now, sigma_now, alpha_next,
                                        sigma_next, _lib.ptr(out), _lib.ptr(gray_out), B, Cc, HW, _lib.stream_of(x))
    return out


_OPS = {"OR": 0, "AND": 1}


def step_ddpm_logq(x, noise_pred, logq, operation, temp, bias, sqrt_one_minus_ab, beta, sqrt_alpha, sqrt_post_var,
                   dtau, z=None, rng=None, out=None, kappa_out=None):
    """src/diffusion/samplers.py:20-58.  ``logq`` [B, K] is updated in place."""
    x, noise_pred, z = _prep(x, noise_pred, z)
    B, Cc, HW = _shape3(x)
    out = torch.empty_like(x) if out is None else out
    op = _OPS.get(str(operation).upper(), 2)
    _call(x, "cdm_step_ddpm_logq", _lib.ptr(x), _lib.ptr_array(noise_pred), len(noise_pred), _lib.ptr(z),
                                             _rng(rng), _lib.ptr(logq), op, temp, bias, sqrt_one_minus_ab, beta,
                                             sqrt_alpha, sqrt_post_var, dtau, _lib.ptr(out), _lib.ptr(kappa_out),
                                             B, Cc, HW, _lib.stream_of(x))
    return out


def step_ode_kappa(x, eps1, eps2, div1, div2, sigma, a, coef, dt, mode=0, div1_scale=1.0, den_eps=1e-9,
                   clip=(-1.0, 2.0), out=None, kappa_out=None):
    """shapes/compose_images_ito.py:66-85,119-135 and the latent variants (see cdm_b200.h)."""
    x, (eps1, eps2), _ = _prep(x, [eps1, eps2], None)
    B, Cc, HW = _shape3(x)
    out = torch.empty_like(x) if out is None else out
    e1c = 1 if x.dim() == 2 else eps1.shape[1]
    _call(x, "cdm_step_ode_kappa", _lib.ptr(x), _lib.ptr(eps1), e1c, _lib.ptr(eps2),
                                             _lib.ptr(div1.float().contiguous()), _lib.ptr(div2.float().contiguous()),
                                             div1_scale, mode, sigma, a, coef, dt, den_eps, clip[0], clip[1],
                                             _lib.ptr(out), _lib.ptr(kappa_out), B, Cc, HW, _lib.stream_of(x))
    return out


def step_ode_kappa_k(x, eps, divs, sigma, a, coef, dt, div_scale=None, den_eps=1e-9, out=None, kappa_out=None):
    """K-expert (2..4) Ito density-ratio weights + probability-flow Euler step in one launch (see cdm_b200.h).  eps[k]:
    [B, 1 or C, ...]; divs[k]: [B]; div_scale[k]: factor on divs[k] (3 for a 1-channel expert repeated over RGB)."""
    x, eps, _ = _prep(x, list(eps), None)
    B, Cc, HW = _shape3(x)
    K = len(eps)
    out = torch.empty_like(x) if out is None else out
    dv = [d.float().contiguous() for d in divs]
    ds = _lib.farray(div_scale) if div_scale is not None else None
    _call(x, "cdm_step_ode_kappa_k", _lib.ptr(x), _lib.ptr_array(eps), _lib.iarray(_channels(eps, x)), _lib.ptr_array(dv), ds, K,
          sigma, a, coef, dt, den_eps, _lib.ptr(out), _lib.ptr(kappa_out), B, Cc, HW, _lib.stream_of(x))
    return out


def step_cfg(x, eps, weights, wsum, combine, update, c0, c1, c2=1.0, c3=0.0, z=None, rng=None, out=None):
    """Guidance-sum (combine 0) / weighted-mean (combine 1) with the x0-form (update 0) or ancestral (update 1) step."""
    x, eps, z = _prep(x, eps, z)
    B, Cc, HW = _shape3(x)
    out = torch.empty_like(x) if out is None else out
    _call(x, "cdm_step_cfg", _lib.ptr(x), _lib.ptr_array(eps), _lib.farray(weights), len(eps), wsum,
                                       combine, update, c0, c1, c2, c3, _lib.ptr(z), _rng(rng), _lib.ptr(out),
                                       B, Cc, HW, _lib.stream_of(x))
    return out


def step_layout(x, eps, masks, masks_f64, s1m, sab, c0, c1, spv, z=None, rng=None, out=None):
    """LayoutDiff step (src/composing_colored_digit_to_simulate_overlaying.py:84-119).  ``masks``: [K, H*W] float64 CUDA
    tensor of per-pixel expert weights; no ``z`` and no ``rng`` = the noise-free last step."""
    x, eps, z = _prep(x, eps, z)
    B, Cc, HW = _shape3(x)
    if masks.dtype != torch.float64 or not masks.is_cuda or masks.shape != (len(eps), HW):
        raise ValueError(f"masks must be a float64 CUDA tensor of shape ({len(eps)}, {HW})")
    out = torch.empty_like(x) if out is None else out
    _call(x, "cdm_step_layout", _lib.ptr(x), _lib.ptr_array(eps), len(eps), _lib.ptr(masks.contiguous()),
                                          1 if masks_f64 else 0, s1m, sab, c0, c1, spv, _lib.ptr(z), _rng(rng), _lib.ptr(out),
                                          B, Cc, HW, _lib.stream_of(x))
    return out


def step_superdiff_solve(x, noise_preds, log_q, mode, temp, bias, som, beta, sqrt_recip_alpha, sqrt_post_var, d_tau, f_coef,
                         g_sq, dw=None, z=None, rng=None, out=None, kappa_out=None):
    """SuperDiff step with the linear-solve kappa (mode "AND", needs ``dw``) or softmax kappa (mode "OR"); K <= 4.
    log_q [B, K] is updated in place.  src/composing_conditional_diffusion_on_shape_and_color_6_1.py:352-428."""
    m = {"OR": 0, "AND": 1}.get(str(mode).upper())
    if m is None:
        raise ValueError("Mode must be 'OR' or 'AND'")
    x, noise_preds, z = _prep(x, noise_preds, z)
    B, Cc, HW = _shape3(x)
    if log_q.shape != (B, len(noise_preds)) or log_q.dtype != torch.float32 or not log_q.is_contiguous():
        raise ValueError("log_q must be a contiguous float32 [B, K] tensor")
    dw = dw.float().contiguous() if dw is not None else None
    out = torch.empty_like(x) if out is None else out
    _call(x, "cdm_step_superdiff_solve", _lib.ptr(x), _lib.ptr_array(noise_preds), len(noise_preds), m, temp, bias, som,
                                                   beta, sqrt_recip_alpha, sqrt_post_var, d_tau, f_coef, g_sq, _lib.ptr(dw),
                                                   _lib.ptr(z), _rng(rng), _lib.ptr(log_q), _lib.ptr(out), _lib.ptr(kappa_out),
                                                   B, Cc, HW, _lib.stream_of(x))
    return out


def decode_latents(latents, components, mean, out=None):
    """PCA inverse transform on the device: ``latents @ components + mean`` (mnist/sample_latent.py:88-89,
    ``pca.inverse_transform`` in shapes/visualize_composition_latent_ito.py:188).  latents [B, L], components [L, D]."""
    _lib.require_cuda(latents)
    dev = latents.device
    z = latents.float().contiguous()
    comp = torch.as_tensor(components, dtype=torch.float32, device=dev).contiguous()
    mu = torch.as_tensor(mean, dtype=torch.float32, device=dev).contiguous()
    B, L = z.shape
    D = comp.shape[1]
    if comp.shape[0] != L or mu.numel() != D:
        raise ValueError(f"components {tuple(comp.shape)} / mean {tuple(mu.shape)} do not match latents {tuple(z.shape)}")
    out = torch.empty(B, D, device=dev, dtype=torch.float32) if out is None else out
    _call(z, "cdm_latent_decode", _lib.ptr(z), _lib.ptr(comp), _lib.ptr(mu), _lib.ptr(out), B, L, D, _lib.stream_of(z))
    return out


def grayscale(x, out=None):
    """torchvision Grayscale(1) of an RGB batch (shapes/compose_images_ddim.py:47)."""
    _lib.require_cuda(x)
    x = x.contiguous()
    B, Cc, HW = _shape3(x)
    if Cc != 3:
        raise ValueError("grayscale expects 3 channels")
    out = torch.empty((B, 1) + tuple(x.shape[2:]), device=x.device, dtype=torch.float32) if out is None else out
    _call(x, "cdm_grayscale", _lib.ptr(x), _lib.ptr(out), B, HW, _lib.stream_of(x))
    return out


def fill_normal(shape, device, rng):
    """The N(0,1) stream the kernels draw for rng=(seed, step), materialised (lets a checker replay it)."""
    z = torch.empty(shape, device=device, dtype=torch.float32)
    _call(z, "cdm_fill_normal", _lib.ptr(z), z.numel(), _rng(rng), _lib.stream_of(z))
    return z
