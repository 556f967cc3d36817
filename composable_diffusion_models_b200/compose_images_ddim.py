"""Two-expert DDIM composition (shape expert on grayscale + colour expert on RGB).

Drop-in for ``shapes/compose_images_ddim.py`` / ``shapes/compose_scores.py``:
``sample_composed_ddim(shape_model, color_model, shape_label, color_label, args)`` with ``args.bs .img_size
.n_steps .w_shape .w_color`` returns x [B, 3, S, S].  Per step: 2 expert forwards on libcdm_b200 + ONE
fused kernel (weighted mean, clamped x0, DDIM update, and the Grayscale of the result for the next step).
"""
import ctypes as C

import torch

from . import _chain, _lib, schedule, steps
from .models import UNet


class Config:
    DEVICE = "cuda"
    SHAPES = ["circle", "square", "triangle"]
    COLORS = ["red", "green", "blue"]


def ddim_tables(n_steps):
    """time grid of shapes/compose_images_ddim.py:37 and alpha/sigma at every grid point (fp32)."""
    ts = torch.linspace(1.0, 1e-3, n_steps + 1)
    return ts, schedule.alpha(ts), schedule.sigma(ts)


def _ddim_chain(models, labels, weights, wsum, x, n_steps):
    """The whole loop in ONE host call (cdm_unet_sample_ddim): per-step scalars are a host table, no Python between kernels."""
    lib = _lib.lib()
    x = x.contiguous()
    B, Cc, S = x.shape[0], x.shape[1], x.shape[2]
    if B == 0:
        return x
    ts, al, sg = ddim_tables(n_steps)
    coef, cptr = _chain.host_coef(torch.stack([ts, al, sg], dim=1))
    prec = _lib.precision_code(models[0].precision)
    harr, hp = _chain.handle_array(models, x.device)
    yp, keep, uniform = _chain.label_arrays(labels, B, x.device)
    with torch.cuda.device(x.device):
        ws = _chain.workspace(x.device, lib.cdm_unet_sample_ddim_workspace_bytes(hp, len(models), B, Cc, S, prec))
        _lib.check(lib.cdm_unet_sample_ddim(hp, _lib.farray(weights), len(models), float(wsum), _lib.ptr(x), yp, uniform, cptr, n_steps,
                                            B, Cc, S, prec, _lib.ptr(ws), ws.numel(), _lib.stream_of(x)))
    del coef, harr, keep
    return x


@torch.no_grad()
def sample_composed_ddim(shape_model, color_model, shape_label, color_label, args, x_init=None, use_chain=None):
    device = Config.DEVICE
    shape_model.eval()
    color_model.eval()
    x = (torch.randn(args.bs, 3, args.img_size, args.img_size, device=device) if x_init is None
         else x_init.to(device).float().clone())
    if use_chain is None:
        use_chain = _chain.native_all([shape_model, color_model], UNet, x) and x.shape[2] == x.shape[3]
    if use_chain:
        return _ddim_chain([shape_model, color_model], [shape_label, color_label], [args.w_shape, args.w_color],
                           args.w_shape + args.w_color, x, args.n_steps)
    ts, al, sg = ddim_tables(args.n_steps)
    ts, al, sg = ts.tolist(), al.tolist(), sg.tolist()
    x_gray = steps.grayscale(x)
    w = [args.w_shape, args.w_color]
    for i in range(args.n_steps):
        t = torch.full((x.shape[0],), ts[i], device=x.device)
        eps_s = shape_model(x_gray, t, shape_label)
        eps_c = color_model(x, t, color_label)
        x = steps.step_ddim(x, [eps_s, eps_c], w, args.w_shape + args.w_color, al[i], sg[i], al[i + 1], sg[i + 1],
                            out=x, gray_out=x_gray)
    return x


@torch.no_grad()
def sample_full_ddim(model, num_samples, num_classes, device, img_size, in_channels, timesteps, labels=None, x_init=None,
                     use_chain=None):
    """K=1 DDIM; ``shapes/train_image.py:43-85``."""
    model.eval()
    x = torch.randn(num_samples, in_channels, img_size, img_size, device=device) if x_init is None else x_init.to(device).float().clone()
    if labels is None and num_classes:
        labels = torch.arange(num_samples, device=device) % num_classes
    if use_chain is None:
        use_chain = _chain.native_all([model], UNet, x) and model.in_channels == in_channels
    if use_chain:
        return _ddim_chain([model], [labels if num_classes else None], [1.0], 1.0, x, timesteps)
    ts, al, sg = ddim_tables(timesteps)
    ts, al, sg = ts.tolist(), al.tolist(), sg.tolist()
    for i in range(timesteps):
        t = torch.full((x.shape[0],), ts[i], device=x.device)
        eps = model(x, t, labels) if num_classes else model(x, t)
        x = steps.step_ddim(x, [eps], [1.0], 1.0, al[i], sg[i], al[i + 1], sg[i + 1], out=x)
    return x
