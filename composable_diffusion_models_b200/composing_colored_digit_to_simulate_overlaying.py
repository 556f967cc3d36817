"""LayoutDiff: spatial-mask composition of K noise-prediction experts.

Drop-in for ``LayoutDiff(sde).sample(models, masks, shape, device)`` and ``create_circular_mask`` of
``src/composing_colored_digit_to_simulate_overlaying.py:56-133``.  Per step the K experts are evaluated and ONE fused
kernel (``cdm_step_layout``) does the masked sum, the clamped-x0 posterior mean and the noise injection; the per-step
scalars come from host copies of the VPSDE tables (no device syncs).  Optional ``x_init=`` / ``noise=`` ([T-1, B, ...])
inject the Gaussian draws; omitted, they are drawn with ``torch.randn`` in the reference's order.
"""
import numpy as np
import torch

from . import steps


def create_circular_mask(h, w, center=None, radius=None):
    """reference :127-133 (returns a float64 tensor, as the reference does)."""
    if center is None:
        center = (int(w / 2), int(h / 2))
    if radius is None:
        radius = min(center[0], center[1], w - center[0], h - center[1])
    Y, X = np.ogrid[:h, :w]
    dist_from_center = np.sqrt((X - center[0]) ** 2 + (Y - center[1]) ** 2)
    return torch.from_numpy((dist_from_center <= radius).astype(float))


def final_masks(masks):
    """Non-overlapping region of every mask; the LAST model in the list is on top (reference :70-79)."""
    out = [torch.zeros_like(m) for m in masks]
    occlusion = torch.zeros_like(masks[0])
    for i in range(len(masks) - 1, -1, -1):
        unique = torch.clamp(masks[i] - occlusion, 0, 1)
        out[i] = unique
        occlusion += unique
    return out


class LayoutDiff:
    """A diffusion sampler that composes models based on spatial masks."""

    def __init__(self, sde):
        self.sde = sde

    @torch.no_grad()
    def sample(self, models, masks, shape, device, x_init=None, noise=None):
        if len(models) != len(masks):
            raise ValueError("The number of models and masks must be equal.")
        x = torch.randn(shape, device=device) if x_init is None else x_init.to(device).float().clone()
        fm = final_masks(list(masks))
        # float64 masks (what create_circular_mask returns) make torch evaluate `eps * mask` in double
        f64 = all(m.dtype == torch.float64 for m in fm)
        mdev = torch.stack([m.reshape(-1).to(torch.float64) for m in fm]).to(device)
        T = self.sde.num_timesteps
        tb = self.sde.host_tables()
        sab_tab = torch.sqrt(tb["alphas_cumprod"])
        for m in models:
            if hasattr(m, "eval"):
                m.eval()
        for i in range(T):
            t_idx = T - 1 - i
            t = torch.full((shape[0],), t_idx, device=device, dtype=torch.long)
            preds = [m(x, t.float()) for m in models]
            ac, abp, beta = tb["alphas_cumprod"][t_idx], tb["alphas_cumprod_prev"][t_idx], tb["betas"][t_idx]
            c0 = float(torch.sqrt(abp) * beta / (1.0 - ac))
            c1 = float(torch.sqrt(tb["alphas"][t_idx]) * (1.0 - abp) / (1.0 - ac))
            s1m = float(tb["sqrt_one_minus_alphas_cumprod"][t_idx])
            sab = float(sab_tab[t_idx])
            if i < T - 1:
                spv = float(torch.sqrt(tb["posterior_variance"][t_idx]))
                z = torch.randn_like(x) if noise is None else noise[i].to(device)
                x = steps.step_layout(x, preds, mdev, f64, s1m, sab, c0, c1, spv, z=z, out=x)
            else:
                x = steps.step_layout(x, preds, mdev, f64, s1m, sab, c0, c1, 0.0, out=x)
        return x.clamp(-1, 1)
