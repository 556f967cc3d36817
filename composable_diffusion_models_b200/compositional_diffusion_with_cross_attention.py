"""Two-condition classifier-free-guidance sampling (x0-form update).

Drop-in for ``sample_composed(Config, model, digit, color_idx)`` of
``src/compositional_diffusion_with_cross_attention.py:266-315``, generalised from batch 1 to
``batch_size`` independent chains.  The reference evaluates 4 label pairs per step and never uses the
fully-conditioned one (:292); only the 3 that enter the update are evaluated here.  As written in the
reference, the guided model output is used both as x0 and as the direction term (:307-313).
"""
import torch

from . import _chain, _lib, steps
from .models import GuidedUNet


def alphas_cumprod(timesteps, beta_start=0.0001, beta_end=0.02):
    """``get_coeff`` (:212-219), host side, once per sampler call instead of once per step."""
    betas = torch.linspace(beta_start, beta_end, timesteps)
    return torch.cumprod(1.0 - betas, axis=0)


@torch.no_grad()
def sample_composed(Config, model, digit, color_idx, batch_size=1, x_init=None, use_chain=None):
    model.eval()
    dev = Config.DEVICE
    x = (torch.randn((batch_size, 3, Config.IMG_SIZE, Config.IMG_SIZE), device=dev) if x_init is None
         else x_init.to(dev).float().clone())
    bs = x.shape[0]
    acp = alphas_cumprod(Config.TIMESTEPS)
    one = torch.tensor(1.0)
    if use_chain is None:
        use_chain = _chain.native_all([model], GuidedUNet, x)
    if use_chain and bs > 0:
        # the whole loop in ONE host call (cdm_guided_sample_cfg); per-step scalars as a host table
        T = Config.TIMESTEPS
        rows = []
        for i in reversed(range(T)):
            ab_prev = acp[i - 1] if i > 0 else one
            rows.append([float(i), float(torch.sqrt(ab_prev)), float(torch.sqrt(1.0 - ab_prev))])
        ctab, cptr = _chain.host_coef(rows)
        lib = _lib.lib()
        x = x.contiguous()
        prec = _lib.precision_code(model.precision)
        h = model._native_handle(x.device)
        with torch.cuda.device(x.device):
            ws = _chain.workspace(x.device, lib.cdm_guided_sample_cfg_workspace_bytes(h, bs, Config.IMG_SIZE, prec))
            _lib.check(lib.cdm_guided_sample_cfg(h, _lib.ptr(x), int(digit), int(color_idx), float(Config.GUIDANCE_STRENGTH_SHAPE),
                                                 float(Config.GUIDANCE_STRENGTH_COLOR), cptr, T, bs, Config.IMG_SIZE, prec,
                                                 _lib.ptr(ws), ws.numel(), _lib.stream_of(x)))
        del ctab
        return (x.clamp(-1, 1) + 1) / 2
    digits = torch.full((bs,), digit, device=dev, dtype=torch.long)
    colors = torch.full((bs,), color_idx, device=dev, dtype=torch.long)
    null_d = torch.full((bs,), model.null_digit_idx, device=dev, dtype=torch.long)
    null_c = torch.full((bs,), model.null_color_idx, device=dev, dtype=torch.long)
    ws, wc = Config.GUIDANCE_STRENGTH_SHAPE, Config.GUIDANCE_STRENGTH_COLOR
    for i in reversed(range(Config.TIMESTEPS)):
        t = torch.full((bs,), i, device=dev)
        p_unc = model(x, t, null_d, null_c)
        p_shape = model(x, t, digits, null_c)
        p_color = model(x, t, null_d, colors)
        ab_prev = acp[i - 1] if i > 0 else one
        c0 = float(torch.sqrt(ab_prev))
        c1 = float(torch.sqrt(1.0 - ab_prev))
        x = steps.step_cfg(x, [p_unc, p_shape, p_color], [1.0, ws, wc], 1.0, 0, 0, c0, c1, out=x)
    return (x.clamp(-1, 1) + 1) / 2
