"""GPU: LayoutDiff spatial-mask composition -- the fused step kernel teacher-forced
against the oracle (bit-exact elementwise), and the whole sampler behind the reference's signature against outputs of the
unmodified reference (tests/golden/sampler_layoutdiff*.npz), fp32 experts, <= 1e-5 rel-L2."""
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import experts as E
from oracle import samplers as OS
from oracle import schedule as S

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _score(seed):
    from composable_diffusion_models_b200.models import ColoredMNISTScoreModel
    m = ColoredMNISTScoreModel(precision="fp32")
    m.load_state_dict(E.synth_state_dict(E.score_model_spec(), seed), strict=True)
    return m.to(DEV).eval()


@pytest.mark.parametrize("f64", [True, False])
@pytest.mark.parametrize("K,B,C,HW", [(2, 3, 3, 32 * 32), (3, 2, 1, 28 * 28), (4, 5, 3, 15 * 15)])
def test_layout_step_bit_exact_vs_oracle(K, B, C, HW, f64):
    from composable_diffusion_models_b200 import steps
    g = torch.Generator().manual_seed(K * 100 + B)
    side = int(HW ** 0.5)
    x = torch.randn(B, C, side, side, generator=g)
    preds = [torch.randn(B, C, side, side, generator=g) for _ in range(K)]
    z = torch.randn(B, C, side, side, generator=g)
    masks = [torch.rand(side, side, generator=g, dtype=torch.float64 if f64 else torch.float32) for _ in range(K)]
    sde = S.VPSDETables(num_timesteps=50)
    fm = OS.layout_final_masks(masks)
    t_idx = 17
    mdev = torch.stack([m.reshape(-1).double() for m in fm]).to(DEV)
    ac, abp, beta = sde.alphas_cumprod[t_idx], sde.alphas_cumprod_prev[t_idx], sde.betas[t_idx]
    c0 = float(torch.sqrt(abp) * beta / (1.0 - ac))
    c1 = float(torch.sqrt(sde.alphas[t_idx]) * (1.0 - abp) / (1.0 - ac))
    s1m, sab = float(sde.sqrt_one_minus_alphas_cumprod[t_idx]), float(sde.sqrt_alphas_cumprod[t_idx])
    spv = float(torch.sqrt(sde.posterior_variance[t_idx]))
    for last in (False, True):
        want = OS.layoutdiff_step(sde, x, preds, fm, t_idx, z, last)
        got = steps.step_layout(x.to(DEV), [p.to(DEV) for p in preds], mdev, f64, s1m, sab, c0, c1, spv,
                                z=None if last else z.to(DEV))
        assert torch.equal(got.cpu(), want.float()), (last, float((got.cpu() - want).abs().max()))


@pytest.mark.parametrize("name,mk", [("sampler_layoutdiff", ("mask_b", "mask_a")), ("sampler_layoutdiff_soft", ("mask0", "mask1"))])
def test_layoutdiff_sampler_vs_reference(name, mk):
    from composable_diffusion_models_b200.composing_colored_digit_to_simulate_overlaying import LayoutDiff
    from composable_diffusion_models_b200.schedule import VPSDE
    g = load_golden(name)
    models = [_score(g["seed1"]), _score(g["seed2"])]
    out = LayoutDiff(VPSDE(num_timesteps=g["T"], device=DEV)).sample(models, [g[mk[0]], g[mk[1]]], (2, 3, 32, 32), DEV,
                                                                  x_init=g["x_init"], noise=g["noise"])
    assert rel_l2(out.cpu(), g["out"]) < 1e-5


def test_layoutdiff_contract():
    from composable_diffusion_models_b200.composing_colored_digit_to_simulate_overlaying import LayoutDiff, create_circular_mask
    from composable_diffusion_models_b200.schedule import VPSDE
    m = create_circular_mask(32, 32)
    assert m.dtype == torch.float64 and m.shape == (32, 32) and m[16, 16] == 1 and m[0, 0] == 0
    with pytest.raises(ValueError):
        LayoutDiff(VPSDE(num_timesteps=4, device=DEV)).sample([_score(1)], [m, m], (1, 3, 32, 32), DEV)
    # unseeded path draws its own noise and stays in range
    out = LayoutDiff(VPSDE(num_timesteps=3, device=DEV)).sample([_score(1), _score(2)], [m, 1 - m], (2, 3, 32, 32), DEV)
    assert out.shape == (2, 3, 32, 32) and out.abs().max() <= 1
