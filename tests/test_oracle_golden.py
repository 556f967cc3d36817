"""CPU: the oracle restatement vs. outputs of the UNMODIFIED reference (tests/golden/*.npz,
made by oracle/make_golden.py).  This is the oracle's pin."""
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import experts as E
from oracle import samplers as OS
from oracle import schedule as S

TOL = 2e-6   # same fp32 CPU ops in (nearly) the same order as the reference


def _check_wsum(sd, want):
    got = float(sum(v.double().abs().sum() for v in sd.values() if v.is_floating_point()))
    assert abs(got - want) <= 1e-9 * abs(want), "synthetic weight generator drifted from the golden fixtures"


def test_schedule_matches_reference():
    g = load_golden("schedule")
    t = g["t"]
    for name, fn in [("log_alpha", S.log_alpha), ("alpha", S.alpha), ("sigma", S.sigma),
                     ("dlog_alphadt", S.dlog_alphadt), ("beta", S.beta), ("g2", S.g2),
                     ("jax_sigma", S.jax_sigma), ("jax_beta", S.jax_beta), ("jax_g2", S.jax_g2)]:
        assert torch.equal(fn(t), g[name]), name
    sde = S.VPSDETables()
    for name in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_one_minus_alphas_cumprod",
                 "posterior_variance"):
        assert torch.equal(getattr(sde, name), g["vpsde_" + name]), name


def test_stable_schedule_is_consistent():
    t = torch.linspace(0.01, 1.0, 17)
    assert torch.allclose(S.stable_sigma(t), S.sigma(t), rtol=1e-4, atol=1e-6)
    assert torch.allclose(S.stable_beta(t), S.beta(t), rtol=2e-4, atol=1e-6)


def test_unet_mnist():
    g = load_golden("unet_mnist")
    sd = E.synth_state_dict(E.unet_small_spec(1), g["seed"])
    _check_wsum(sd, g["wsum"])
    assert len(sd) == 66          # strict=True load into the reference module pinned this key set
    assert rel_l2(E.unet_small_forward(sd, g["x"], g["t"]), g["eps"]) < TOL


def test_unet_shapes_conditional():
    g = load_golden("unet_shapes")
    sds = E.synth_state_dict(E.unet_small_spec(1, num_classes=3), g["seed_shape"])
    sdc = E.synth_state_dict(E.unet_small_spec(3, num_classes=3), g["seed_color"])
    _check_wsum(sds, g["wsum_shape"])
    _check_wsum(sdc, g["wsum_color"])
    assert rel_l2(E.unet_small_forward(sds, g["x_shape"], g["t"], g["y"]), g["eps_shape"]) < TOL
    assert rel_l2(E.unet_small_forward(sdc, g["x_color"], g["t"], g["y"]), g["eps_color"]) < TOL
    with pytest.raises(ValueError):
        E.unet_small_forward(sds, g["x_shape"], g["t"], None)


def test_mlp_2d():
    g = load_golden("mlp_2d")
    sd = E.synth_state_dict(E.mlp_2d_spec(), g["seed"])
    _check_wsum(sd, g["wsum"])
    assert rel_l2(E.mlp_2d_forward(sd, g["t"], g["x"]), g["eps"]) < TOL


def test_score_model():
    g = load_golden("score_model")
    sd = E.synth_state_dict(E.score_model_spec(), g["seed"])
    _check_wsum(sd, g["wsum"])
    assert rel_l2(E.score_model_forward(sd, g["x"], g["t"]), g["eps"]) < TOL


def test_guided_unet_and_degenerate_attention():
    g = load_golden("guided_unet")
    sd = E.synth_state_dict(E.guided_unet_spec(), g["seed"])
    _check_wsum(sd, g["wsum"])
    assert rel_l2(E.guided_unet_forward(sd, g["x"], g["t"], g["digits"], g["colors"]), g["eps"]) < 5e-6


def test_sampler_sde_mnist():
    g = load_golden("sampler_sde_mnist")
    sd1 = E.synth_state_dict(E.unet_small_spec(1), g["seed1"])
    sd2 = E.synth_state_dict(E.unet_small_spec(1), g["seed2"])
    ex = [lambda x, t: E.unet_small_forward(sd1, x, t), lambda x, t: E.unet_small_forward(sd2, x, t)]
    out = OS.sample_sde(ex, [g["w1"], g["w2"]], g["x_init"], g["noise"], g["n_steps"], g["xi"])
    assert rel_l2(out, g["out"]) < TOL


def _shape_color(g):
    sds = E.synth_state_dict(E.unet_small_spec(1, num_classes=3), g["seed_shape"])
    sdc = E.synth_state_dict(E.unet_small_spec(3, num_classes=3), g["seed_color"])

    def fs(x, t):
        return E.unet_small_forward(sds, x, t, torch.full((x.shape[0],), g["shape_label"], dtype=torch.long))

    def fc(x, t):
        return E.unet_small_forward(sdc, x, t, torch.full((x.shape[0],), g["color_label"], dtype=torch.long))
    return fs, fc


def test_sampler_ddim():
    g = load_golden("sampler_ddim")
    fs, fc = _shape_color(g)
    out = OS.sample_ddim(fs, fc, g["x_init"], g["n_steps"], g["w_shape"], g["w_color"])
    assert rel_l2(out, g["out"]) < TOL


@pytest.mark.parametrize("variant", ["beta", "g2"])
def test_sampler_ito(variant):
    g = load_golden(f"sampler_ito_{variant}")
    fs, fc = _shape_color(g)
    probes = list(zip(g["probes_shape"], g["probes_color"]))
    out = OS.sample_ito_ode(fs, fc, g["x_init"], probes, g["n_steps"], variant)
    assert rel_l2(out, g["out"]) < 2e-5     # kappa divides by sum((s1-s2)^2): ill-conditioned


@pytest.mark.parametrize("op", ["or", "and", "avg"])
def test_sampler_superdiff(op):
    g = load_golden(f"sampler_superdiff_{op}")
    sd1 = E.synth_state_dict(E.score_model_spec(), g["seed1"])
    sd2 = E.synth_state_dict(E.score_model_spec(), g["seed2"])
    ex = [lambda x, t: E.score_model_forward(sd1, x, t), lambda x, t: E.score_model_forward(sd2, x, t)]
    sde = S.VPSDETables(num_timesteps=g["T"])
    out, _ = OS.sample_superdiff(sde, ex, g["x_init"], g["noise"], op.upper(), g["temp"], 0.0)
    assert rel_l2(out, g["out"]) < TOL


def test_sampler_ddpm_single():
    g = load_golden("sampler_ddpm_single")
    sd = E.synth_state_dict(E.score_model_spec(), g["seed"])
    sde = S.VPSDETables(num_timesteps=g["T"])
    out = OS.sample_ddpm_single(sde, lambda x, t: E.score_model_forward(sd, x, t), g["x_init"], g["noise"])
    assert rel_l2(out, g["out"]) < TOL


def test_sampler_cfg_x0():
    g = load_golden("sampler_cfg_x0")
    sd = E.synth_state_dict(E.guided_unet_spec(), g["seed"])
    out = OS.sample_cfg_x0(lambda x, t, d, c: E.guided_unet_forward(sd, x, t, d, c), g["x_init"],
                           g["digit"], g["color"], 10, 3, g["timesteps"])
    assert rel_l2(out, g["out"]) < 5e-6


@pytest.mark.parametrize("name,mk", [("sampler_layoutdiff", ("mask_b", "mask_a")), ("sampler_layoutdiff_soft", ("mask0", "mask1"))])
def test_sampler_layoutdiff(name, mk):
    """section 8(f) row 1: binary float64 circle masks (the reference's own experiment) and soft float32 masks."""
    g = load_golden(name)
    sd1 = E.synth_state_dict(E.score_model_spec(), g["seed1"])
    sd2 = E.synth_state_dict(E.score_model_spec(), g["seed2"])
    ex = [lambda x, t: E.score_model_forward(sd1, x, t), lambda x, t: E.score_model_forward(sd2, x, t)]
    sde = S.VPSDETables(num_timesteps=g["T"])
    out = OS.sample_layoutdiff(sde, ex, [g[mk[0]], g[mk[1]]], g["x_init"], g["noise"])
    assert rel_l2(out, g["out"]) < TOL


@pytest.mark.parametrize("name", ["sampler_superdiff61_and_l0", "sampler_superdiff61_and_l3", "sampler_superdiff61_or_l0"])
def test_sampler_superdiff_linear_solve(name):
    """section 8(f) row 2: the oracle's K-expert restatement run at K = 2 against the unmodified reference
    (src/composing_conditional_diffusion_on_shape_and_color_6_1.py sample_superdiff, batch 1)."""
    g = load_golden(name)
    sds = [E.synth_state_dict(E.score_model_spec(), s) for s in (g["seed1"], g["seed2"])]
    ex = [lambda x, t, sd=sd: E.score_model_forward(sd, x, t.float()) for sd in sds]
    mode = "AND" if "_and_" in name else "OR"
    out, _ = OS.sample_superdiff_6_1(g["T"], ex, g["x_init"], g["dw"], g["noise"], mode, g["temp"], g["bias"])
    assert rel_l2(out, g["out"]) < TOL


def test_beta_vae_decode():
    """section 8(f) row 3: the oracle's decoder vs the unmodified reference class (src/4.3 best_of_both_worlds_3.py BetaVAE)."""
    g = load_golden("beta_vae_decode")
    sd = E.synth_state_dict(E.beta_vae_spec(g["latent_dims"]), g["seed"])
    assert rel_l2(E.beta_vae_decode(sd, g["z"]), g["out"]) < TOL


def test_save_image_quantize():
    """the quantisation restatement vs pixels written by torchvision.utils.save_image itself (bit-exact)."""
    g = load_golden("save_image_quantize")
    assert torch.equal(E.save_image_quantize(g["x"].clone()).permute(1, 2, 0), g["u8_hwc"])


def test_simple_unet_forward():
    """section 8(f) row 4: the oracle's SimpleUnet vs the unmodified reference class (..._shape_and_color_6.py SimpleUnet)."""
    g = load_golden("simple_unet")
    sd = E.synth_state_dict(E.simple_unet_spec(g["num_classes"]), g["seed"])
    assert rel_l2(E.simple_unet_forward(sd, g["x"], g["t"], g["y"]), g["out"]) < TOL


def test_diffusion_sde_3_tables_and_sampler():
    """row a2: the DiffusionSDE tables (finite-difference f_t, g_t^2, div f_t) and the batched sample_superdiff of
    src/composing_conditional_diffusion_on_shape_and_color_3.py -- the oracle AND the product's host-side class against the
    unmodified reference (tables bit-equal; sampler: see oracle/make_golden_3.py for the one neutralised expression)."""
    from composable_diffusion_models_b200.composing_conditional_diffusion_on_shape_and_color_3 import DiffusionSDE
    g = load_golden("diffusion_sde_3_tables")
    names = ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
             "posterior_variance", "f_t_coeff", "g_t_sq", "div_f_t")
    for T in (500, 8):
        tb = OS.diffusion_sde_3_tables(T, (3, 32, 32))
        d = DiffusionSDE(T, (3, 32, 32), "cpu")
        for k in names:
            assert torch.equal(torch.as_tensor(tb[k]).float(), g[f"T{T}_{k}"].float()), (T, k)
            assert torch.equal(getattr(d, k).float(), g[f"T{T}_{k}"].float()), (T, k)
    for strategy in ("or", "avg"):
        g = load_golden(f"sampler_superdiff3_{strategy}")
        sds = [E.synth_state_dict(E.score_model_spec(), s) for s in (g["seed1"], g["seed2"])]
        ex = [lambda x, t, sd=sd: E.score_model_forward(sd, x, t.float()) for sd in sds]
        out, _ = OS.sample_superdiff_3(g["T"], ex, g["x_init"], g["noise"], strategy.upper(), g["temp"], g["bias"])
        assert rel_l2(out, g["out"]) < TOL


@pytest.mark.parametrize("name,variant", [("latent_ito", "stable"), ("latent_ito_2", "clipped")])
def test_latent_ito_script_loops(name, variant):
    """The sampling loops of shapes/visualize_composition_latent_ito.py:117-147 and _ito_2.py:93-119 -- top-level scripts,
    pinned by executing their own loop source under stub globals (oracle/make_golden_latent.py)."""
    g = load_golden(name)
    sds = [E.synth_state_dict(E.mlp_2d_spec(), int(g[k])) for k in ("seed1", "seed2")]
    n, x = int(g["n_steps"]), g["x_init"].clone()
    dt = 1.0 / n
    for i in range(n):
        t_val = 1.0 - i * dt
        t = torch.full((x.shape[0],), t_val)
        e1, d1 = E.hutchinson_vjp_div(lambda xx: E.mlp_2d_forward(sds[0], t, xx), x, g["probes"][i, 0])
        e2, d2 = E.hutchinson_vjp_div(lambda xx: E.mlp_2d_forward(sds[1], t, xx), x, g["probes"][i, 1])
        x, _ = OS.latent_ito_step(x, e1, e2, d1, d2, t_val, dt, variant)
    # kappa = num / (den + 1e-9) is ill-conditioned where the two experts agree: the loop amplifies last-bit differences
    assert rel_l2(x, g["out"]) < 1e-5


def test_latent_sde_script_loop():
    """The sampling loop of mnist/visualize_composition_latent.py:63-87 (weighted-sum reverse SDE on 2-D latents)."""
    g = load_golden("latent_sde")
    sds = [E.synth_state_dict(E.mlp_2d_spec(), int(g[k])) for k in ("seed1", "seed2")]
    fns = [lambda x, t, sd=sd: E.mlp_2d_forward(sd, t, x) for sd in sds]
    out = OS.sample_sde(fns, [float(g["w1"]), float(g["w2"])], g["x_init"], g["noise"], int(g["n_steps"]), 1.0)
    assert rel_l2(out, g["out"]) < TOL
