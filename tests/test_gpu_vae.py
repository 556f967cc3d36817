"""GPU: the BetaVAE decoder (the latent samplers' image-space epilogue) and the save_image
quantisation on libcdm_b200, vs the reference-generated goldens and the oracle at other sizes."""
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import experts as E

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-5


def _vae(latent, seed):
    from composable_diffusion_models_b200.models import BetaVAE
    m = BetaVAE(latent)
    sd = E.synth_state_dict(E.beta_vae_spec(latent), seed)
    m.load_state_dict(sd, strict=True)       # the full BetaVAE state_dict (encoder half included) loads
    return m.to(DEV).eval(), sd


def test_decode_vs_reference_golden():
    g = load_golden("beta_vae_decode")
    m, _ = _vae(g["latent_dims"], g["seed"])
    got = m.decode(g["z"].to(DEV))
    assert got.shape == g["out"].shape
    assert rel_l2(got.cpu(), g["out"]) < TOL


@pytest.mark.parametrize("latent,B", [(10, 1), (10, 37), (4, 300), (32, 9)])
def test_decode_vs_oracle(latent, B):
    m, sd = _vae(latent, 150 + latent)
    z = torch.randn(B, latent, generator=torch.Generator().manual_seed(B)) * 2
    got = m.decode(z.to(DEV))
    want = E.beta_vae_decode(sd, z)
    assert rel_l2(got.cpu(), want) < TOL
    assert float(got.min()) > 0.0 and float(got.max()) < 1.0


def test_decode_edge_cases():
    m, _ = _vae(10, 7)
    assert m.decode(torch.empty(0, 10, device=DEV)).shape == (0, 3, 32, 32)
    with pytest.raises(ValueError):
        m.decode(torch.zeros(2, 11, device=DEV))
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 3, 32, 32, device=DEV))


def test_latent_chain_then_decode_matches_oracle():
    """The use the reference makes of it: decode the final latents of a sampling chain (sample_composed_latent, :262-293)."""
    m, sd = _vae(10, 9)
    z = torch.randn(16, 10, generator=torch.Generator().manual_seed(1))
    a = torch.linspace(0.99, 0.9, 20)
    for i in range(20):                       # any deterministic latent trajectory will do for the decoder's parity
        z = z / a[i].sqrt() - 0.05 * torch.tanh(z)
    from composable_diffusion_models_b200.models import quantize_u8
    img = m.decode(z.to(DEV))
    assert rel_l2(img.cpu(), E.beta_vae_decode(sd, z)) < TOL
    # quantised pixels: identical except where the float image sits within rounding distance of a .5 boundary
    q_gpu, q_ref = quantize_u8(img).cpu(), E.save_image_quantize(E.beta_vae_decode(sd, z))
    assert (q_gpu.int() - q_ref.int()).abs().max() <= 1 and (q_gpu != q_ref).float().mean() < 1e-3


def test_quantize_bit_exact():
    g = load_golden("save_image_quantize")
    from composable_diffusion_models_b200.models import quantize_u8
    got = quantize_u8(g["x"].to(DEV)).cpu()
    assert got.dtype == torch.uint8 and torch.equal(got.permute(1, 2, 0), g["u8_hwc"])
    x = torch.rand(5, 3, 33, 17, generator=torch.Generator().manual_seed(3)) * 1.4 - 0.2
    assert torch.equal(quantize_u8(x.to(DEV)).cpu(), E.save_image_quantize(x.clone()))
    assert quantize_u8(torch.empty(0, 3, 4, 4, device=DEV)).numel() == 0
