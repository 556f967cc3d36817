"""GPU: SuperDiff with the linear-solve kappa -- cdm_step_superdiff_solve teacher-forced
against the oracle's K-expert restatement (K = 2, 3, 4; interior and clamped kappas; singular systems), and the sampler
behind the reference's signature against outputs of the unmodified reference (K = 2, batch 1)."""
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import experts as E
from oracle import samplers as OS

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _step_inputs(K, B, C, S, seed, spread):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, S, S, generator=g)
    base = torch.randn(B, C, S, S, generator=g)
    preds = [base + spread * torch.randn(B, C, S, S, generator=g) for _ in range(K)]
    dw = torch.randn(B, C, S, S, generator=g)
    z = torch.randn(B, C, S, S, generator=g)
    log_q = torch.randn(B, K, generator=g)
    return x, preds, dw, z, log_q


def _gpu_step(tb, T, i, x, preds, log_q, dw, z, mode, temp, bias):
    from composable_diffusion_models_b200 import steps
    from composable_diffusion_models_b200.superdiff_linear_solve import get_forward_process_params
    f_coef, g_sq = get_forward_process_params(tb, i, T)
    lq = log_q.clone().to(DEV)
    kap = torch.zeros_like(lq)
    out = steps.step_superdiff_solve(x.to(DEV), [p.to(DEV) for p in preds], lq, mode, temp, bias,
                                     float(tb["sqrt_one_minus_alphas_cumprod"][i]), float(tb["betas"][i]),
                                     float(tb["sqrt_recip_alphas"][i]), float(torch.sqrt(tb["posterior_variance"][i])) if i > 0 else 0.0,
                                     1.0 / T, f_coef, g_sq, dw=dw.to(DEV) if mode == "AND" else None,
                                     z=z.to(DEV) if i > 0 else None, kappa_out=kap)
    return out.cpu(), lq.cpu(), kap.cpu()


@pytest.mark.parametrize("mode", ["AND", "OR"])
@pytest.mark.parametrize("K,B,C,S,spread", [(2, 3, 3, 32, 1.0), (2, 2, 3, 32, 0.05), (3, 4, 1, 28, 1.0), (4, 2, 3, 16, 0.5), (4, 3, 3, 15, 1.0),
                                               (2, 3, 3, 64, 1.0), (4, 2, 3, 64, 0.5)])   # 64x64: cluster-split samples
def test_superdiff_solve_step_vs_oracle(K, B, C, S, spread, mode):
    T = 50
    tb = OS.ddpm_tables_6_1(T)
    for i, bias in ((37, 0.0), (5, 0.4), (0, 0.0)):
        x, preds, dw, z, log_q = _step_inputs(K, B, C, S, 100 * K + i, spread)
        want_x, want_q, want_k = OS.superdiff_6_1_step(tb, x, preds, log_q, i, dw, z, mode, 1.3, bias)
        got_x, got_q, got_k = _gpu_step(tb, T, i, x, preds, log_q, dw, z, mode, 1.3, bias)
        # kappa comes out of a (possibly ill-conditioned) K x K solve built from differently-ordered fp32 reductions
        assert (got_k - want_k).abs().max() < 2e-3, (i, got_k, want_k)
        assert abs(float(got_k.sum(1).mean()) - 1.0) < 1e-5
        assert rel_l2(got_x, want_x) < 2e-4, i
        assert rel_l2(got_q, want_q) < 1e-4, i


def test_superdiff_solve_singular_system_gives_uniform_kappa():
    """Identical experts make rows of the system identical (LinAlgError in the reference -> (0.5, 0.5))."""
    T = 20
    tb = OS.ddpm_tables_6_1(T)
    x, preds, dw, z, log_q = _step_inputs(2, 2, 3, 16, 7, 1.0)
    preds = [preds[0], preds[0].clone()]
    _, _, kap = _gpu_step(tb, T, 9, x, preds, log_q, dw, z, "AND", 1.0, 0.0)
    assert torch.allclose(kap, torch.full_like(kap, 0.5))


@pytest.mark.parametrize("name", ["sampler_superdiff61_and_l0", "sampler_superdiff61_and_l3", "sampler_superdiff61_or_l0"])
def test_sample_superdiff_vs_reference(name):
    import types
    from composable_diffusion_models_b200.models import ColoredMNISTScoreModel
    from composable_diffusion_models_b200.superdiff_linear_solve import sample_superdiff
    g = load_golden(name)
    ms = []
    for seed in (g["seed1"], g["seed2"]):
        m = ColoredMNISTScoreModel(precision="fp32")
        m.load_state_dict(E.synth_state_dict(E.score_model_spec(), seed), strict=True)
        m = m.to(DEV).eval()
        ms.append(lambda img, t, lab, m=m: m(img, t.float()))
    mode = "AND" if "_and_" in name else "OR"
    cfg = types.SimpleNamespace(DEVICE=DEV, IMG_SIZE=32, TIMESTEPS=g["T"])
    out = sample_superdiff(ms[0], ms[1], 0, 1, mode=mode, T=g["temp"], l=g["bias"], x_init=g["x_init"], dw=g["dw"],
                           noise=g["noise"], config=cfg)
    assert rel_l2(out.cpu(), g["out"]) < 1e-4
    with pytest.raises(ValueError):
        sample_superdiff(ms[0], ms[1], 0, 1, mode="XOR", config=cfg)


@pytest.mark.parametrize("strategy", ["or", "avg"])
def test_sample_superdiff_3_vs_reference(strategy):
    """row a2: DiffusionSDE + the batched sample_superdiff of src/composing_conditional_diffusion_on_shape_and_color_3.py
    (kappa-weighted noise, p_sample, log-density update with the finite-difference f_t / g_t^2 tables) against the
    unmodified reference's output, and its log-densities against the oracle."""
    from composable_diffusion_models_b200 import composing_conditional_diffusion_on_shape_and_color_3 as M3
    from composable_diffusion_models_b200.models import ColoredMNISTScoreModel
    g = load_golden(f"sampler_superdiff3_{strategy}")
    ms, sds = [], []
    for seed in (g["seed1"], g["seed2"]):
        sd = E.synth_state_dict(E.score_model_spec(), seed)
        m = ColoredMNISTScoreModel(precision="fp32")
        m.load_state_dict(sd, strict=True)
        m = m.to(DEV).eval()
        ms.append(lambda x, t, c, m=m: m(x, t.float()))
        sds.append(sd)
    M3.Config.IMG_SIZE = 32
    diffusion = M3.DiffusionSDE(g["T"], (3, 32, 32), DEV)
    out, lq = M3.sample_superdiff(ms[0], ms[1], diffusion, 0, 1, num_images=3, strategy=strategy.upper(), temp=g["temp"],
                                  bias=g["bias"], x_init=g["x_init"], noise=g["noise"], return_log_q=True)
    assert rel_l2(out.cpu(), g["out"]) < 1e-5
    _, want_lq = OS.sample_superdiff_3(g["T"], [lambda x, t, sd=sd: E.score_model_forward(sd, x, t.float()) for sd in sds],
                                       g["x_init"], g["noise"], strategy.upper(), g["temp"], g["bias"])
    assert rel_l2(lq.cpu(), want_lq) < 1e-4
    # p_sample / q_sample keep the reference's signatures
    x = g["x_init"].to(DEV)
    t = torch.full((3,), 5, device=DEV, dtype=torch.long)
    e = torch.randn(3, 3, 32, 32, generator=torch.Generator().manual_seed(1)).to(DEV)
    z = torch.randn(3, 3, 32, 32, generator=torch.Generator().manual_seed(2)).to(DEV)
    h = diffusion.host_tables()
    want = torch.sqrt(1.0 / h["alphas"])[5] * (x.cpu() - h["betas"][5] * e.cpu() / h["sqrt_one_minus_alphas_cumprod"][5]) \
        + torch.sqrt(h["posterior_variance"][5]) * z.cpu()
    assert rel_l2(diffusion.p_sample(e, x, t, noise=z).cpu(), want) < 1e-6
    assert diffusion.q_sample(x, t, noise=z).shape == x.shape
