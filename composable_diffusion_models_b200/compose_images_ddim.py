"""Two-expert DDIM composition (shape expert on grayscale + colour expert on RGB).

Drop-in for ``shapes/compose_images_ddim.py`` / ``shapes/compose_scores.py``:
``sample_composed_ddim(shape_model, color_model, shape_label, color_label, args)`` with ``args.bs .img_size
.n_steps .w_shape .w_color`` returns x [B, 3, S, S].  Per step: 2 expert forwards on libcdm_b200 + ONE
fused kernel (weighted mean, clamped x0, DDIM update, and the Grayscale of the result for the next step).
"""
import torch

from . import schedule, steps


class Config:
    DEVICE = "cuda"
    SHAPES = ["circle", "square", "triangle"]
    COLORS = ["red", "green", "blue"]


def ddim_tables(n_steps):
    """time grid of shapes/compose_images_ddim.py:37 and alpha/sigma at every grid point (fp32)."""
    ts = torch.linspace(1.0, 1e-3, n_steps + 1)
    return ts, schedule.alpha(ts), schedule.sigma(ts)


@torch.no_grad()
def sample_composed_ddim(shape_model, color_model, shape_label, color_label, args, x_init=None):
    device = Config.DEVICE
    shape_model.eval()
    color_model.eval()
    x = (torch.randn(args.bs, 3, args.img_size, args.img_size, device=device) if x_init is None
         else x_init.to(device).float().clone())
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
def sample_full_ddim(model, num_samples, num_classes, device, img_size, in_channels, timesteps, labels=None, x_init=None):
    """K=1 DDIM; ``shapes/train_image.py:43-85``."""
    model.eval()
    x = torch.randn(num_samples, in_channels, img_size, img_size, device=device) if x_init is None else x_init.to(device).float().clone()
    if labels is None and num_classes:
        labels = torch.arange(num_samples, device=device) % num_classes
    ts, al, sg = ddim_tables(timesteps)
    ts, al, sg = ts.tolist(), al.tolist(), sg.tolist()
    for i in range(timesteps):
        t = torch.full((x.shape[0],), ts[i], device=x.device)
        eps = model(x, t, labels) if num_classes else model(x, t)
        x = steps.step_ddim(x, [eps], [1.0], 1.0, al[i], sg[i], al[i + 1], sg[i + 1], out=x)
    return x
