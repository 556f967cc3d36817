"""GPU: (1) the K-expert Ito density-ratio step (BASELINE config 4 names four experts) against the oracle's K-expert
semantics, which IS the reference's get_kappa at K = 2; (2) the whole-chain C entries (cdm_unet_sample_ddim,
cdm_unet_sample_ito, cdm_score_sample_superdiff, cdm_guided_sample_cfg) against the per-step loops they replace -- same
kernels, same per-step scalars, so the results are bit-identical."""
import types

import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import experts as E
from oracle import samplers as OS
from oracle import schedule as S

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _unet(kw, seed, precision):
    from composable_diffusion_models_b200.models import UNet
    m = UNet(**kw, precision=precision)
    m.load_state_dict(E.synth_state_dict(E.unet_small_spec(kw.get("in_channels", 1), num_classes=kw.get("num_classes")), seed), strict=True)
    return m.to(DEV).eval()


# ---------------------------------------------------------------------------------------------------------------------
# K-expert kappa step
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("K,chs", [(2, (1, 3)), (3, (1, 3, 3)), (4, (1, 1, 3, 3)), (4, (3, 3, 3, 3)), (3, (3, 1, 3))])
@pytest.mark.parametrize("B,size", [(5, 16), (3, 64), (2, 7)])
def test_ode_kappa_k_step_vs_oracle(K, chs, B, size):
    from composable_diffusion_models_b200 import steps
    g = torch.Generator().manual_seed(K * 100 + B + size)
    x = torch.randn(B, 3, size, size, generator=g)
    eps = [torch.randn(B, c, size, size, generator=g) for c in chs]
    divs = [torch.randn(B, generator=g) * 20 for _ in chs]
    scale = [3.0 if c == 1 else 1.0 for c in chs]
    t_val, dt = 0.63, 1e-2
    tt = torch.tensor(t_val)
    sig, a, coef = float(S.sigma(tt)), float(S.dlog_alphadt(tt)), 0.5 * float(S.beta(tt))
    eps_rgb = [e.repeat(1, 3, 1, 1) if e.shape[1] == 1 else e for e in eps]
    want, kap = OS.ito_ode_step_k(x, eps_rgb, [d * s for d, s in zip(divs, scale)], t_val, dt, "beta")
    kout = torch.zeros(B, K, device=DEV)
    got = steps.step_ode_kappa_k(x.to(DEV), [e.to(DEV) for e in eps], [d.to(DEV) for d in divs], sig, a, coef, dt, div_scale=scale,
                                 kappa_out=kout)
    assert rel_l2(got.cpu(), want) < 2e-6
    if K == 2:      # routed to the two-expert kernel: bit-identical to cdm_step_ode_kappa
        two = steps.step_ode_kappa(x.to(DEV), eps[0].to(DEV), eps[1].to(DEV), divs[0].to(DEV), divs[1].to(DEV), sig, a, coef, dt,
                                   mode=0, div1_scale=3.0)
        assert torch.equal(got, two)
    else:
        assert rel_l2(kout.cpu(), kap) < 1e-4
        assert torch.allclose(kout.sum(1).cpu(), torch.ones(B), atol=1e-5)


def test_ode_kappa_k_equal_density_rates():
    """The defining property, checked on the kernel's own kappa in double precision: every expert's log-density changes at
    the same rate along the composed flow."""
    from composable_diffusion_models_b200 import steps
    g = torch.Generator().manual_seed(4)
    B, S, K = 6, 16, 4
    x = torch.randn(B, 3, S, S, generator=g)
    eps = [torch.randn(B, 3, S, S, generator=g) for _ in range(K)]
    divs = [torch.randn(B, generator=g) * 50 for _ in range(K)]
    sig = 0.8
    kout = torch.zeros(B, K, device=DEV)
    steps.step_ode_kappa_k(x.to(DEV), [e.to(DEV) for e in eps], [d.to(DEV) for d in divs], sig, -3.0, 2.0, 1e-3, kappa_out=kout, den_eps=0.0)
    kap = kout.cpu().double()
    s = [-(e.double()) / sig for e in eps]
    dv = [-(d.double()) / sig for d in divs]
    sc = sum(kap[:, j].view(-1, 1, 1, 1) * s[j] for j in range(K))
    rates = torch.stack([dv[k] - (s[k] * (sc - s[k])).sum(dim=(1, 2, 3)) for k in range(K)], dim=1)
    spread = (rates.max(1).values - rates.min(1).values) / rates.abs().mean(1)
    assert spread.max() < 1e-3, spread


def test_ito_k4_sampler_vs_oracle():
    """Four shapes experts (two shape-type + two colour-type), a few probability-flow steps, injected probes; fp32 path."""
    from composable_diffusion_models_b200 import compose_images_ito as ITO
    chs, seeds, labs = (1, 1, 3, 3), (611, 612, 613, 614), (2, 0, 1, 2)
    B, S_, n = 2, 16, 4
    models = [_unet(dict(in_channels=c, num_classes=3), s, "fp32") for c, s in zip(chs, seeds)]
    sds = [E.synth_state_dict(E.unet_small_spec(c, num_classes=3), s) for c, s in zip(chs, seeds)]
    g = torch.Generator().manual_seed(8)
    x0 = torch.randn(B, 3, S_, S_, generator=g)
    probes = [[torch.randn(B, c, S_, S_, generator=g) for c in chs] for _ in range(n)]
    labels_h = [torch.full((B,), v, dtype=torch.long) for v in labs]
    want = OS.sample_ito_ode_k([lambda x, t, sd=sd, y=y: E.unet_small_forward(sd, x, t, y) for sd, y in zip(sds, labels_h)], chs, x0,
                               probes, n, "beta")
    args = types.SimpleNamespace(bs=B, img_size=S_, n_steps=n)
    labels = [y.to(DEV) for y in labels_h]
    loop = ITO.sample_composed_ito_ode_k(models, labels, args, x_init=x0, probes=probes, use_chain=False)
    chain = ITO.sample_composed_ito_ode_k(models, labels, args, x_init=x0, probes=probes)
    assert rel_l2(loop.cpu(), want) < 2e-4          # the fp32 JVP divergence vs autograd: see test_gpu_ito.py
    assert torch.equal(chain, loop)


# ---------------------------------------------------------------------------------------------------------------------
# whole-chain entries == per-step loops
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "f16x3", "fp16"])
def test_ddim_chain_entry_is_the_per_step_loop(precision):
    from composable_diffusion_models_b200 import compose_images_ddim as D
    ms = _unet(dict(in_channels=1, num_classes=3), 321, precision)
    mc = _unet(dict(in_channels=3, num_classes=3), 322, precision)
    B, S_, n = 5, 32, 9
    x0 = torch.randn(B, 3, S_, S_, generator=torch.Generator().manual_seed(2))
    sl = torch.full((B,), 2, dtype=torch.long, device=DEV)
    cl = torch.tensor([0, 1, 2, 1, 0], device=DEV)            # non-uniform labels: B embedding rows in the chain too
    args = types.SimpleNamespace(bs=B, img_size=S_, n_steps=n, w_shape=1.5, w_color=0.5)
    loop = D.sample_composed_ddim(ms, mc, sl, cl, args, x_init=x0, use_chain=False)
    chain = D.sample_composed_ddim(ms, mc, sl, cl, args, x_init=x0)
    assert torch.equal(chain, loop)
    # uniform labels: the chain embeds ONE row per step and expert, the loop B rows -- same summation order, same bits
    cl2 = torch.full((B,), 1, dtype=torch.long, device=DEV)
    assert torch.equal(D.sample_composed_ddim(ms, mc, sl, cl2, args, x_init=x0), D.sample_composed_ddim(ms, mc, sl, cl2, args, x_init=x0, use_chain=False))
    # K = 1 (shapes/train_image.py sample_full_ddim)
    a = D.sample_full_ddim(mc, B, 3, DEV, S_, 3, n, x_init=x0)
    b = D.sample_full_ddim(mc, B, 3, DEV, S_, 3, n, x_init=x0, use_chain=False)
    assert torch.equal(a, b)


@pytest.mark.parametrize("variant", ["beta", "g2"])
@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_ito_chain_entry_is_the_per_step_loop(variant, precision):
    from composable_diffusion_models_b200 import compose_images_ito as ITO
    ms = _unet(dict(in_channels=1, num_classes=3), 331, precision)
    mc = _unet(dict(in_channels=3, num_classes=3), 332, precision)
    B, S_, n = 3, 16, 5
    g = torch.Generator().manual_seed(3)
    x0 = torch.randn(B, 3, S_, S_, generator=g)
    c0 = 1 if variant == "beta" else 3
    probes = [(torch.randn(B, c0, S_, S_, generator=g), torch.randn(B, 3, S_, S_, generator=g)) for _ in range(n)]
    sl, cl = torch.full((B,), 2, dtype=torch.long, device=DEV), torch.full((B,), 1, dtype=torch.long, device=DEV)
    args = types.SimpleNamespace(bs=B, img_size=S_, n_steps=n)
    loop = ITO.sample_composed_ito_ode(ms, mc, sl, cl, args, variant=variant, x_init=x0, probes=probes, use_chain=False)
    chain = ITO.sample_composed_ito_ode(ms, mc, sl, cl, args, variant=variant, x_init=x0, probes=probes)
    if variant == "beta":
        assert torch.equal(chain, loop)
    else:       # the loop sums the probe's channels with torch, the chain with its own kernel: same values, maybe not the same order
        assert rel_l2(chain.cpu(), loop.cpu()) < (1e-6 if precision == "fp32" else 2e-3)
    # torch-drawn probes (the reference's RNG order) and library-drawn probes both run and stay finite
    torch.manual_seed(5)
    a = ITO.sample_composed_ito_ode(ms, mc, sl, cl, args, variant=variant, x_init=x0)
    torch.manual_seed(5)
    b = ITO.sample_composed_ito_ode(ms, mc, sl, cl, args, variant=variant, x_init=x0, use_chain=False)
    assert rel_l2(a.cpu(), b.cpu()) < (1e-6 if precision == "fp32" else 2e-3)
    c = ITO.sample_composed_ito_ode(ms, mc, sl, cl, args, variant=variant, x_init=x0, seed=11)
    d = ITO.sample_composed_ito_ode(ms, mc, sl, cl, args, variant=variant, x_init=x0, seed=11)
    assert torch.equal(c, d) and torch.isfinite(c).all()


@pytest.mark.parametrize("op", ["OR", "AND", "AVG"])
def test_superdiff_chain_entry_is_the_per_step_loop(op):
    from composable_diffusion_models_b200.diffusion import SuperDiffSampler
    from composable_diffusion_models_b200.models import ColoredMNISTScoreModel
    from composable_diffusion_models_b200.schedule import VPSDE
    experts = []
    for seed in (41, 42, 43):
        m = ColoredMNISTScoreModel()
        m.load_state_dict(E.synth_state_dict(E.score_model_spec(), seed), strict=True)
        experts.append(m.to(DEV).eval())
    T, B = 7, 4
    g = torch.Generator().manual_seed(6)
    x0 = torch.randn(B, 3, 32, 32, generator=g)
    noise = torch.randn(T, B, 3, 32, 32, generator=g)
    sampler = SuperDiffSampler(VPSDE(num_timesteps=T, device=DEV))
    kw = dict(operation=op, temp=1.3, bias=0.1, x_init=x0, noise=noise, return_log_q=True)
    for models in (experts[:2], experts):
        loop, lq1 = sampler.sample(models[0], models[1], B, (3, 32, 32), DEV, models=models, use_chain=False, **kw)
        chain, lq2 = sampler.sample(models[0], models[1], B, (3, 32, 32), DEV, models=models, **kw)
        assert torch.equal(chain, loop) and torch.equal(lq1, lq2)
    a = sampler.sample(experts[0], experts[1], B, (3, 32, 32), DEV, operation=op, x_init=x0, noise="kernel", seed=3)
    b = sampler.sample(experts[0], experts[1], B, (3, 32, 32), DEV, operation=op, x_init=x0, noise="kernel", seed=3, use_chain=False)
    assert torch.equal(a, b)


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_cfg_chain_entry_is_the_per_step_loop(precision):
    from composable_diffusion_models_b200.compositional_diffusion_with_cross_attention import sample_composed
    from composable_diffusion_models_b200.models import GuidedUNet
    m = GuidedUNet(precision=precision)
    m.load_state_dict(E.synth_state_dict(E.guided_unet_spec(), 9), strict=True)
    m = m.to(DEV).eval()
    cfg = types.SimpleNamespace(DEVICE=DEV, IMG_SIZE=32, TIMESTEPS=6, GUIDANCE_STRENGTH_SHAPE=7.5, GUIDANCE_STRENGTH_COLOR=5.0)
    x0 = torch.randn(3, 3, 32, 32, generator=torch.Generator().manual_seed(1))
    loop = sample_composed(cfg, m, 7, 2, batch_size=3, x_init=x0, use_chain=False)
    chain = sample_composed(cfg, m, 7, 2, batch_size=3, x_init=x0)
    assert torch.equal(chain, loop)


# ---------------------------------------------------------------------------------------------------------------------
# grouped K-expert launches (cdm_unet_forward_grouped)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("K,B,S", [(2, 5, 28), (3, 33, 28), (4, 2, 28), (2, 3, 64), (2, 130, 28), (2, 1, 32)])
def test_grouped_forward_is_k_separate_forwards(K, B, S):
    """One grouped launch per convolution (gridDim.y = expert) against K separate forwards: the same kernel bodies,
    parameters and per-tile arithmetic, order-independent statistics -> bit-identical."""
    from composable_diffusion_models_b200.models import forward_grouped
    experts = [_unet(dict(in_channels=1), 700 + k, "fp16") for k in range(K)]
    g = torch.Generator().manual_seed(K + B)
    x = torch.randn(B, 1, S, S, generator=g).to(DEV)
    t = (torch.rand(B, generator=g) * 0.9 + 0.05).to(DEV)
    want = [m(x, t) for m in experts]
    got = forward_grouped(experts, x, t)
    for a, b in zip(got, want):
        assert torch.equal(a, b)


def test_grouped_forward_mixed_channel_experts_and_fallbacks():
    from composable_diffusion_models_b200 import _lib, steps
    from composable_diffusion_models_b200.models import forward_grouped
    ms = _unet(dict(in_channels=1, num_classes=3), 801, "fp16")
    mc = _unet(dict(in_channels=3, num_classes=3), 802, "fp16")
    g = torch.Generator().manual_seed(1)
    B = 6
    x = torch.randn(B, 3, 32, 32, generator=g).to(DEV)
    xg = steps.grayscale(x)
    t = torch.full((B,), 0.7, device=DEV)
    ys = [torch.randint(0, 3, (B,), generator=g).to(DEV), torch.randint(0, 3, (B,), generator=g).to(DEV)]
    got = forward_grouped([ms, mc], [xg, x], t, ys)
    assert torch.equal(got[0], ms(xg, t, ys[0])) and torch.equal(got[1], mc(x, t, ys[1]))
    # fp32-class experts and a single expert are not grouped: the entry point runs them back to back, same results
    m32 = [_unet(dict(in_channels=1), 810 + k, "fp32") for k in range(2)]
    x1 = torch.randn(3, 1, 28, 28, generator=g).to(DEV)
    t1 = torch.full((3,), 0.3, device=DEV)
    for a, m in zip(forward_grouped(m32, x1, t1), m32):
        assert torch.equal(a, m(x1, t1))
    lib = _lib.lib()
    try:        # the switch the chain entries consult
        lib.cdm_set_option(b"grouped", 0)
        a = forward_grouped([ms, mc], [xg, x], t, ys)
    finally:
        lib.cdm_set_option(b"grouped", -1)
    assert torch.equal(a[0], got[0]) and torch.equal(a[1], got[1])
    with pytest.raises(ValueError):
        forward_grouped([ms, mc], [xg, x], t)          # conditional experts need labels
