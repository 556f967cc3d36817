"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Test infrastructure only.  Run in the build container (``/root/reference`` is
not present on the GPU box; the committed fixtures are what travels):

    python -m oracle.make_golden

What it does: imports the reference's own modules from ``/root/reference`` by
path, loads seeded synthetic weights (``oracle.experts.synth_state_dict``) with
``load_state_dict(strict=True)`` -- which also pins the checkpoint key set --
and runs the reference's sampler functions with ``torch.randn`` /
``torch.randn_like`` replaced by a recording, seeded generator so the exact
noise the reference consumed can be replayed into the oracle and the CUDA path.
Modules the reference needs but this image lacks (matplotlib, imageio) are
stubbed in ``sys.modules``; no reference source is modified or copied.
"""
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np
import torch

from . import experts as E

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


class NoiseTap:
    """Replaces torch.randn / torch.randn_like with a seeded, recording source."""

    def __init__(self, seed):
        self.g = torch.Generator().manual_seed(seed)
        self.draws = []
        self._orig = (torch.randn, torch.randn_like)

    def randn(self, *size, **kw):
        if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
            size = tuple(size[0])
        z = self._orig[0](*size, generator=self.g, dtype=torch.float32)
        self.draws.append(z.clone())
        return z

    def randn_like(self, x, **kw):
        return self.randn(*x.shape)

    def __enter__(self):
        torch.randn, torch.randn_like = self.randn, self.randn_like
        return self

    def __exit__(self, *a):
        torch.randn, torch.randn_like = self._orig


def _wsum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values() if v.is_floating_point()))


def _save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    conv = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        conv[k] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **conv)
    print("wrote", name, {k: tuple(v.shape) for k, v in conv.items()})


def main():
    torch.set_num_threads(8)
    torch.manual_seed(0)
    for m in ("matplotlib", "matplotlib.pyplot", "imageio", "box"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["box"].Box = dict

    mnist_sched = _load(f"{REF}/mnist/schedule.py", "ref_mnist_schedule")
    sched2 = _load(f"{REF}/shapes/schedule_2.py", "ref_schedule_2")
    schedj = _load(f"{REF}/shapes/schedule_jax_faithful.py", "ref_schedule_jax")
    mnist_unet = _load(f"{REF}/mnist/models/unet_small.py", "ref_mnist_unet")
    shapes_unet = _load(f"{REF}/shapes/models/unet_small.py", "ref_shapes_unet")
    mlp_mod = _load(f"{REF}/mnist/models/mlp_2d.py", "ref_mlp")
    score_mod = _load(f"{REF}/src/models/compose_grayscale_object_and_color.py", "ref_score_model")
    superdiff_mod = _load(f"{REF}/src/diffusion/samplers.py", "ref_superdiff")

    # ---- schedules (a1, a2) -------------------------------------------------
    t = torch.cat([torch.linspace(1e-3, 1.0, 41), torch.tensor([1.0 - i / 1000 for i in (0, 1, 500, 998, 999)])])
    sde = score_mod.VPSDE()
    _save("schedule",
          t=t, log_alpha=mnist_sched.log_alpha(t), alpha=mnist_sched.alpha(t), sigma=mnist_sched.sigma(t),
          dlog_alphadt=mnist_sched.dlog_alphadt(t), beta=mnist_sched.beta(t), g2=sched2.g2(t),
          jax_sigma=schedj.sigma(t), jax_beta=schedj.beta(t), jax_g2=schedj.g2(t),
          vpsde_betas=sde.betas, vpsde_alphas_cumprod=sde.alphas_cumprod,
          vpsde_alphas_cumprod_prev=sde.alphas_cumprod_prev,
          vpsde_sqrt_one_minus_alphas_cumprod=sde.sqrt_one_minus_alphas_cumprod,
          vpsde_posterior_variance=sde.posterior_variance)

    g = torch.Generator().manual_seed(7)

    # ---- experts (a3-a8) ----------------------------------------------------
    sd_m = E.synth_state_dict(E.unet_small_spec(1), 101)
    m = mnist_unet.UNet().eval()
    m.load_state_dict(sd_m, strict=True)
    x = torch.randn(3, 1, 28, 28, generator=g)
    tt = torch.tensor([1.0, 0.5, 0.013])
    with torch.no_grad():
        _save("unet_mnist", seed=101, wsum=_wsum(sd_m), x=x, t=tt, eps=m(x, tt))

    sd_s = E.synth_state_dict(E.unet_small_spec(1, num_classes=3), 102)
    sd_c = E.synth_state_dict(E.unet_small_spec(3, num_classes=3), 103)
    ms = shapes_unet.UNet(in_channels=1, num_classes=3).eval()
    mc = shapes_unet.UNet(in_channels=3, num_classes=3).eval()
    ms.load_state_dict(sd_s, strict=True)
    mc.load_state_dict(sd_c, strict=True)
    xs = torch.randn(2, 1, 32, 32, generator=g)
    xc = torch.randn(2, 3, 32, 32, generator=g)
    ys = torch.tensor([2, 0])
    t2 = torch.tensor([0.9, 0.2])
    with torch.no_grad():
        _save("unet_shapes", seed_shape=102, seed_color=103, wsum_shape=_wsum(sd_s), wsum_color=_wsum(sd_c),
              x_shape=xs, x_color=xc, y=ys, t=t2, eps_shape=ms(xs, t2, ys), eps_color=mc(xc, t2, ys))

    sd_mlp = E.synth_state_dict(E.mlp_2d_spec(), 104)
    mm = mlp_mod.MLP().eval()
    mm.load_state_dict(sd_mlp, strict=True)
    xl = torch.randn(16, 2, generator=g)
    tl = torch.rand(16, generator=g)
    with torch.no_grad():
        _save("mlp_2d", seed=104, wsum=_wsum(sd_mlp), x=xl, t=tl, eps=mm(tl, xl))

    sd_sc = E.synth_state_dict(E.score_model_spec(), 105)
    sm = score_mod.ColoredMNISTScoreModel().eval()
    sm.load_state_dict(sd_sc, strict=True)
    x32 = torch.randn(2, 3, 32, 32, generator=g)
    t32 = torch.tensor([999.0, 3.0])
    with torch.no_grad():
        _save("score_model", seed=105, wsum=_wsum(sd_sc), x=x32, t=t32, eps=sm(x32, t32))

    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)                      # some reference modules mkdir relative to cwd at import
    sys.path[:0] = [REF, f"{REF}/shapes"]
    try:
        xattn = _load(f"{REF}/src/compositional_diffusion_with_cross_attention.py", "ref_xattn")
        sd_g = E.synth_state_dict(E.guided_unet_spec(), 106)
        gm = xattn.GuidedUNet().eval()
        gm.load_state_dict(sd_g, strict=True)
        tg = torch.tensor([499, 17])
        dg = torch.tensor([7, 10])
        cg = torch.tensor([3, 1])
        with torch.no_grad():
            _save("guided_unet", seed=106, wsum=_wsum(sd_g), x=x32, t=tg, digits=dg, colors=cg,
                  eps=gm(x32, tg, dg, cg))

        # ---- a13: CFG x0-form sampler (reference batch 1) ------------------
        cfg = xattn.Configuration("g", "r", tmp, device="cpu")
        cfg.TIMESTEPS = 6
        with NoiseTap(11) as tap, torch.no_grad():
            out = xattn.sample_composed(cfg, gm, 7, 2)
        _save("sampler_cfg_x0", seed=106, timesteps=6, digit=7, color=2, x_init=tap.draws[0], out=out)

        # ---- a10: two-expert DDIM ------------------------------------------
        ddim = _load(f"{REF}/shapes/compose_images_ddim.py", "ref_ddim")
        ddim.Config.DEVICE = "cpu"
        args = types.SimpleNamespace(bs=2, img_size=32, n_steps=6, w_shape=1.0, w_color=0.6)
        sl = torch.full((2,), 2, dtype=torch.long)
        cl = torch.full((2,), 1, dtype=torch.long)
        with NoiseTap(12) as tap:
            out = ddim.sample_composed_ddim(ms, mc, sl, cl, args)
        _save("sampler_ddim", seed_shape=102, seed_color=103, n_steps=6, w_shape=1.0, w_color=0.6,
              shape_label=2, color_label=1, x_init=tap.draws[0], out=out)

        # ---- a11: Ito ODE, both variants -----------------------------------
        for variant, fname in (("beta", "compose_images_ito.py"), ("g2", "compose_images_ito_2.py")):
            ito = _load(f"{REF}/shapes/{fname}", f"ref_ito_{variant}")
            ito.Config.DEVICE = "cpu"
            args = types.SimpleNamespace(bs=2, img_size=16, n_steps=4)
            with NoiseTap(13) as tap:
                out = ito.sample_composed_ito_ode(ms, mc, sl, cl, args)
            probes = tap.draws[1:]
            _save(f"sampler_ito_{variant}", seed_shape=102, seed_color=103, n_steps=4, shape_label=2,
                  color_label=1, x_init=tap.draws[0],
                  probes_shape=torch.stack(probes[0::2]), probes_color=torch.stack(probes[1::2]), out=out)
    finally:
        os.chdir(cwd)

    # ---- a12: SuperDiff ---------------------------------------------------------
    sd_sc2 = E.synth_state_dict(E.score_model_spec(), 107)
    sm2 = score_mod.ColoredMNISTScoreModel().eval()
    sm2.load_state_dict(sd_sc2, strict=True)
    sde8 = score_mod.VPSDE(num_timesteps=8)
    sampler = superdiff_mod.SuperDiffSampler(sde8)
    for op in ("OR", "AND", "AVG"):
        with NoiseTap(14) as tap:
            out = sampler.sample(sm, sm2, 2, (3, 32, 32), "cpu", operation=op, temp=1.5, bias=0.0)
        _save(f"sampler_superdiff_{op.lower()}", seed1=105, seed2=107, T=8, temp=1.5,
              x_init=tap.draws[0], noise=torch.stack(tap.draws[1:]), out=out)
    with NoiseTap(15) as tap:
        out = sampler.sample_single_model(sm, 2, (3, 32, 32), "cpu")
    _save("sampler_ddpm_single", seed=105, T=8, x_init=tap.draws[0], noise=torch.stack(tap.draws[1:]), out=out)

    # ---- a9: mnist/compose_scores.main, run for real with stubbed viz ----------
    tmp2 = tempfile.mkdtemp()
    sd_m2 = E.synth_state_dict(E.unet_small_spec(1), 108)
    for sd_, nm in ((sd_m, "e1.pth"), (sd_m2, "e2.pth")):
        torch.save({"epoch": 0, "model_state_dict": sd_, "optimizer_state_dict": {}}, os.path.join(tmp2, nm))
    captured = {}
    viz = types.ModuleType("viz")
    viz.save_grid = lambda x, path, **kw: captured.__setitem__("x", x.clone())
    sys.modules["viz"] = viz
    saved_path = list(sys.path)
    for k in ("models", "models.unet_small", "schedule", "utils"):
        sys.modules.pop(k, None)
    sys.path[:0] = [f"{REF}/mnist"]
    try:
        cs = _load(f"{REF}/mnist/compose_scores.py", "ref_mnist_compose_scores")
        args = types.SimpleNamespace(model1_path=os.path.join(tmp2, "e1.pth"), model2_path=os.path.join(tmp2, "e2.pth"),
                                     output_file=os.path.join(tmp2, "out", "grid.png"), w1=1.0, w2=0.7, bs=2,
                                     n_steps=5, xi=0.8)
        with NoiseTap(16) as tap:
            cs.main(args)
    finally:
        sys.path[:] = saved_path
    _save("sampler_sde_mnist", seed1=101, seed2=108, n_steps=5, w1=1.0, w2=0.7, xi=0.8,
          x_init=tap.draws[0], noise=torch.stack(tap.draws[1:]), out=captured["x"])


if __name__ == "__main__":
    main()
