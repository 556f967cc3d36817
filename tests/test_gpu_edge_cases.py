"""GPU: edge cases of the drop-in boundary -- empty and single-sample batches, non-contiguous inputs, odd sizes, and the
error behaviour the reference has (ValueError for a conditional UNet without labels, ...) or that this library adds
(no CPU fallback, unsupported shapes fail loudly)."""
import pytest
import torch

from conftest import rel_l2
from oracle import experts as E

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _unet(kw, seed, precision):
    from composable_diffusion_models_b200.models import UNet
    m = UNet(**kw, precision=precision)
    sd = E.synth_state_dict(E.unet_small_spec(kw.get("in_channels", 1), num_classes=kw.get("num_classes")), seed)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval(), sd


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_empty_batch_everywhere(precision):
    from composable_diffusion_models_b200 import steps
    from composable_diffusion_models_b200.compose_scores import sample_composed_latent_sde
    from composable_diffusion_models_b200.models import MLP, GuidedUNet
    m, _ = _unet(dict(in_channels=1), 3, precision)
    out = m(torch.zeros(0, 1, 28, 28, device=DEV), torch.zeros(0, device=DEV))
    assert out.shape == (0, 1, 28, 28)
    x = torch.zeros(0, 1, 28, 28, device=DEV)
    assert steps.step_sde(x, [x, x], [1.0, 1.0], 0.1, 0.2, 1e-3, 0.3, z=x).shape == (0, 1, 28, 28)
    g = GuidedUNet(precision=precision).to(DEV).eval()
    e = torch.zeros(0, dtype=torch.long, device=DEV)
    assert g(torch.zeros(0, 3, 32, 32, device=DEV), torch.zeros(0, device=DEV), e, e).shape == (0, 3, 32, 32)
    mlps = [MLP().to(DEV).eval() for _ in range(2)]
    assert sample_composed_latent_sde(mlps, [1.0, 1.0], 0, 5, noise="kernel", seed=1, precision=precision).shape == (0, 2)
    assert steps.decode_latents(torch.zeros(0, 2, device=DEV), torch.zeros(2, 784), torch.zeros(784)).shape == (0, 784)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("f16x3", 1e-5), ("fp16", 2e-3)])
@pytest.mark.parametrize("cin,S", [(1, 28), (3, 64), (1, 12), (3, 20)])
def test_single_sample_and_odd_sizes(cin, S, precision, tol):
    """B = 1, and image sizes that are multiples of 4 but of nothing else (12, 20): every conv falls back to whichever
    kernel supports the map, results unchanged."""
    nc = 3 if cin == 3 else None
    m, sd = _unet(dict(in_channels=cin, num_classes=nc), 77 + S, precision)
    g = torch.Generator().manual_seed(S)
    x = torch.randn(1, cin, S, S, generator=g)
    t = torch.tensor([0.42])
    y = torch.tensor([1]) if nc else None
    want = E.unet_small_forward(sd, x, t, y)
    got = m(x.to(DEV), t.to(DEV), y.to(DEV) if nc else None).cpu()
    assert rel_l2(got, want) < tol


def test_non_contiguous_and_broadcast_inputs():
    m, sd = _unet(dict(in_channels=3, num_classes=3), 5, "fp32")
    g = torch.Generator().manual_seed(1)
    xt = torch.randn(3, 32, 32, 4, generator=g)
    x = xt.permute(3, 0, 1, 2)                    # [4, 3, 32, 32], not contiguous
    assert not x.is_contiguous()
    t = torch.tensor(0.3)                         # scalar time, broadcast over the batch like the reference's schedule calls
    y = torch.tensor([0, 1, 2, 1])
    want = E.unet_small_forward(sd, x.contiguous(), t.expand(4), y)
    got = m(x.to(DEV), t.to(DEV).expand(4), y.to(DEV)).cpu()
    assert rel_l2(got, want) < 1e-5


def test_error_behaviour():
    from composable_diffusion_models_b200 import _lib, steps
    from composable_diffusion_models_b200.models import UNet
    m, _ = _unet(dict(in_channels=3, num_classes=3), 5, "fp32")
    x = torch.zeros(2, 3, 32, 32, device=DEV)
    t = torch.zeros(2, device=DEV)
    with pytest.raises(ValueError):               # shapes/models/unet_small.py:99-101
        m(x, t)
    with pytest.raises(NotImplementedError):      # img_size must be a multiple of 4 (two 2x2 max-pools)
        m(torch.zeros(1, 3, 30, 30, device=DEV), t[:1], torch.zeros(1, dtype=torch.long, device=DEV))
    with pytest.raises(ValueError):
        UNet(precision="bf16").to(DEV)(torch.zeros(1, 1, 28, 28, device=DEV), t[:1])
    with pytest.raises((RuntimeError, ValueError, TypeError)):   # no CPU fallback: host tensors are refused
        m(x.cpu(), t.cpu(), torch.zeros(2, dtype=torch.long))
    with pytest.raises(ValueError):               # more experts than the fused kernels are built for
        steps.step_sde(x, [x] * (_lib.MAX_EXPERTS + 1), [1.0] * (_lib.MAX_EXPERTS + 1), 0.1, 0.2, 1e-3, 0.3, z=x)
    with pytest.raises(ValueError):               # expert output of the wrong shape
        steps.step_sde(x, [x[:, :2]], [1.0], 0.1, 0.2, 1e-3, 0.3, z=x)
