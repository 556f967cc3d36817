"""GPU: Hutchinson divergence by forward-mode differentiation (cdm_unet_forward_jvp / cdm_mlp_forward_jvp) and the
Ito/kappa samplers vs the oracle's autograd VJP and the reference's golden outputs (fp32 path)."""
import types

import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import experts as E
from oracle import samplers as OS

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _unet(kw, seed, precision="fp32"):
    from composable_diffusion_models_b200.models import UNet
    m = UNet(**kw, precision=precision)
    sd = E.synth_state_dict(E.unet_small_spec(kw.get("in_channels", 1), num_classes=kw.get("num_classes")), seed)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval(), sd


@pytest.mark.parametrize("precision,tol_eps,tol_div", [("fp32", 1e-5, 1e-4), ("fp16", 2e-3, 1e-1)])
@pytest.mark.parametrize("cin,S", [(1, 16), (3, 16), (1, 28), (3, 64)])
def test_unet_jvp_matches_autograd_vjp(cin, S, precision, tol_eps, tol_div):
    """fp16 mode: primal and tangent convs on the tensor cores.  The divergence v^T J v is a sum of D signed terms (heavy
    cancellation), so its relative error is ~10x the forward's; it is a one-probe Hutchinson estimate whose own sampling
    noise is O(1).  Measured 0.3-3.4e-2: besides rounding, a 2x2 max-pool window whose two largest fp16 values tie routes
    its tangent through a different element than fp32 does; on 16x16 inputs one such flip moves a sample's estimate by ~1
    on values of ~15 (measured up to 8.9e-2 on the u != v form; bound 1e-1; 64x64: < 2e-2).

    The same discontinuity exists in fp32: GroupNorm statistics are accumulated with float atomics, so activations vary
    by ~5e-6 from run to run, and a pooling window whose two largest values are closer than that flips its tangent
    route between runs (tools/jvp_determinism.py: the 64x64 inputs of seeds 3, 5, 7 have such a window and their
    divergence is bimodal at the 2e-3 level; seeds 4, 6, 8 have none and repeat to 2e-6).  The 64x64 case therefore
    uses a tie-free input."""
    nc = 3 if S in (16, 64) else None
    m, sd = _unet(dict(in_channels=cin, num_classes=nc), 900 + cin, precision)
    g = torch.Generator().manual_seed(4 if S == 64 else 3)
    # fp16 on 16x16 maps: one pooling-tie flip moves ONE sample's estimate by ~1 on values of ~15, so with 3 samples the
    # relative L2 error is a coin toss around the bound (measured 0.09-0.12 depending on which window flips); 12 samples
    # make the same event a 2-3e-2 effect and the test measures the kernels rather than one tie
    B = 12 if (precision != "fp32" and S == 16) else 3
    x = torch.randn(B, cin, S, S, generator=g)
    v = torch.randn(B, cin, S, S, generator=g)
    t = torch.rand(B, generator=g) * 0.9 + 0.05
    y = torch.randint(0, 3, (B,), generator=g) if nc else None
    want_eps, want_div = E.hutchinson_vjp_div(lambda xx: E.unet_small_forward(sd, xx, t, y), x, v)
    eps, div = m.forward_jvp(x.to(DEV), t.to(DEV), y.to(DEV) if nc else None, v.to(DEV))
    assert rel_l2(eps.cpu(), want_eps) < tol_eps
    assert rel_l2(div.cpu(), want_div) < tol_div
    # bilinear form with distinct tangent / cotangent (the "divergence through Grayscale" case)
    u = torch.randn(B, cin, S, S, generator=g)
    with torch.enable_grad():
        xc = x.clone().requires_grad_(True)
        out = E.unet_small_forward(sd, xc, t, y)
        uj = torch.autograd.grad(out, xc, grad_outputs=u)[0]
    want = (uj * v).flatten(1).sum(1)
    _, got = m.forward_jvp(x.to(DEV), t.to(DEV), y.to(DEV) if nc else None, v.to(DEV), u.to(DEV))
    assert rel_l2(got.cpu(), want) < tol_div


def test_mlp_jvp_matches_autograd_vjp():
    from composable_diffusion_models_b200.models import MLP
    sd = E.synth_state_dict(E.mlp_2d_spec(), 51)
    m = MLP()
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV)
    g = torch.Generator().manual_seed(2)
    B = 77
    x, v, t = torch.randn(B, 2, generator=g), torch.randn(B, 2, generator=g), torch.rand(B, generator=g)
    want_eps, want_div = E.hutchinson_vjp_div(lambda xx: E.mlp_2d_forward(sd, t, xx), x, v)
    eps, div = m.forward_jvp(t.to(DEV), x.to(DEV), v.to(DEV))
    assert rel_l2(eps.cpu(), want_eps) < 1e-5
    assert rel_l2(div.cpu(), want_div) < 1e-4


@pytest.mark.parametrize("variant", ["beta", "g2"])
def test_sample_composed_ito_ode_vs_reference(variant):
    from composable_diffusion_models_b200 import compose_images_ito as I
    g = load_golden(f"sampler_ito_{variant}")
    ms, _ = _unet(dict(in_channels=1, num_classes=3), g["seed_shape"])
    mc, _ = _unet(dict(in_channels=3, num_classes=3), g["seed_color"])
    args = types.SimpleNamespace(bs=2, img_size=16, n_steps=g["n_steps"])
    sl = torch.full((2,), g["shape_label"], dtype=torch.long, device=DEV)
    cl = torch.full((2,), g["color_label"], dtype=torch.long, device=DEV)
    probes = list(zip(g["probes_shape"], g["probes_color"]))
    out = I.sample_composed_ito_ode(ms, mc, sl, cl, args, variant=variant, x_init=g["x_init"], probes=probes)
    # kappa = num / (sum (s1-s2)^2 + 1e-9) is ill-conditioned and 4 steps of dt = 0.25 amplify it: 1e-4
    assert rel_l2(out.cpu(), g["out"]) < 1e-4


@pytest.mark.parametrize("variant", ["stable", "clipped"])
def test_sample_latent_ito_ode_vs_oracle(variant):
    from composable_diffusion_models_b200.compose_images_ito import sample_latent_ito_ode
    from composable_diffusion_models_b200.models import MLP
    sds = [E.synth_state_dict(E.mlp_2d_spec(), s) for s in (61, 62)]
    ms = []
    for sd in sds:
        m = MLP()
        m.load_state_dict(sd, strict=True)
        ms.append(m.to(DEV))
    g = torch.Generator().manual_seed(8)
    B, n_steps = 50, 30
    x0 = torch.randn(B, 2, generator=g)
    probes = [(torch.randn(B, 2, generator=g), torch.randn(B, 2, generator=g)) for _ in range(n_steps)]
    x = x0.clone()
    dt = 1.0 / n_steps
    for i in range(n_steps):
        t_val = 1.0 - i * dt
        t = torch.full((B,), t_val)
        e1, d1 = E.hutchinson_vjp_div(lambda xx: E.mlp_2d_forward(sds[0], t, xx), x, probes[i][0])
        e2, d2 = E.hutchinson_vjp_div(lambda xx: E.mlp_2d_forward(sds[1], t, xx), x, probes[i][1])
        x, _ = OS.latent_ito_step(x, e1, e2, d1, d2, t_val, dt, variant)
    got = sample_latent_ito_ode(ms[0], ms[1], B, n_steps, variant=variant, device=DEV, x_init=x0, probes=probes)
    assert rel_l2(got.cpu(), x) < 1e-4


@pytest.mark.parametrize("variant", ["beta", "g2"])
def test_sample_composed_ito_ode_fp16_tracks_fp32(variant):
    """The whole Ito kappa-ODE sampler with primal and tangent convs on the tensor cores (fp16), same probes, against the
    reference's golden output (outputs of the unmodified reference).  Measured 4.2e-4 / 4.8e-4 rel-L2 on the final
    samples; bound 5e-3."""
    from composable_diffusion_models_b200 import compose_images_ito as I
    g = load_golden(f"sampler_ito_{variant}")
    ms, _ = _unet(dict(in_channels=1, num_classes=3), g["seed_shape"], "fp16")
    mc, _ = _unet(dict(in_channels=3, num_classes=3), g["seed_color"], "fp16")
    args = types.SimpleNamespace(bs=2, img_size=16, n_steps=g["n_steps"])
    sl = torch.full((2,), g["shape_label"], dtype=torch.long, device=DEV)
    cl = torch.full((2,), g["color_label"], dtype=torch.long, device=DEV)
    probes = list(zip(g["probes_shape"], g["probes_color"]))
    out = I.sample_composed_ito_ode(ms, mc, sl, cl, args, variant=variant, x_init=g["x_init"], probes=probes)
    err = rel_l2(out.cpu(), g["out"])
    assert err < 5e-3, err


@pytest.mark.parametrize("name,variant", [("latent_ito", "stable"), ("latent_ito_2", "clipped")])
def test_sample_latent_ito_ode_vs_reference_script(name, variant):
    """The 2-D latent Ito samplers against the output of the reference SCRIPTS' own loops
    (shapes/visualize_composition_latent_ito.py:117-147, _ito_2.py:93-119; fixtures by oracle/make_golden_latent.py)."""
    from composable_diffusion_models_b200.compose_images_ito import sample_latent_ito_ode
    from composable_diffusion_models_b200.models import MLP
    g = load_golden(name)
    ms = []
    for k in ("seed1", "seed2"):
        m = MLP()
        m.load_state_dict(E.synth_state_dict(E.mlp_2d_spec(), int(g[k])), strict=True)
        ms.append(m.to(DEV))
    n = int(g["n_steps"])
    probes = [(g["probes"][i, 0], g["probes"][i, 1]) for i in range(n)]
    got = sample_latent_ito_ode(ms[0], ms[1], g["x_init"].shape[0], n, variant=variant, device=DEV, x_init=g["x_init"], probes=probes)
    assert rel_l2(got.cpu(), g["out"]) < 1e-4


def test_latent_sde_vs_reference_script():
    """mnist/visualize_composition_latent.py:63-87 (the reference script's own loop) on the fp32 latent sampler."""
    from composable_diffusion_models_b200.compose_scores import sample_composed_latent_sde
    from composable_diffusion_models_b200.models import MLP
    g = load_golden("latent_sde")
    ms = []
    for k in ("seed1", "seed2"):
        m = MLP()
        m.load_state_dict(E.synth_state_dict(E.mlp_2d_spec(), int(g[k])), strict=True)
        ms.append(m.to(DEV))
    n = int(g["n_steps"])
    got = sample_composed_latent_sde(ms, [float(g["w1"]), float(g["w2"])], g["x_init"].shape[0], n, 1.0, device=DEV,
                                     x_init=g["x_init"], noise=g["noise"])
    assert rel_l2(got.cpu(), g["out"]) < 1e-5
