"""GPU: run-to-run bit-identity.  GroupNorm statistics are accumulated by many CTAs with atomics; they are kept in
64-bit fixed point (csrc/layers.cuh) so the sums -- and every output after them -- do not depend on the order the
atomics land in.  A seeded reference run is bit-reproducible; so is this path."""
import pytest
import torch

from oracle import experts as E

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _unet(kw, seed, precision):
    from composable_diffusion_models_b200.models import UNet
    m = UNet(**kw, precision=precision)
    m.load_state_dict(E.synth_state_dict(E.unet_small_spec(kw.get("in_channels", 1), num_classes=kw.get("num_classes")), seed), strict=True)
    return m.to(DEV).eval()


@pytest.mark.parametrize("precision", ["fp16", "f16x3", "fp32"])
@pytest.mark.parametrize("cin,S,B", [(1, 28, 300), (3, 64, 24), (1, 32, 65)])
def test_forward_twice_bit_identical(precision, cin, S, B):
    from composable_diffusion_models_b200 import _lib
    if precision == "f16x3" and not hasattr(_lib, "PREC_F16X3"):
        pytest.skip("f16x3 mode not built")
    if precision == "fp32" and B > 100:
        B = 40
    nc = 3 if cin == 3 else None
    m = _unet(dict(in_channels=cin, num_classes=nc), 77, precision)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, cin, S, S, generator=g).to(DEV)
    t = (torch.rand(B, generator=g) * 0.9 + 0.05).to(DEV)
    y = torch.randint(0, 3, (B,), generator=g).to(DEV) if nc else None
    first = m(x, t, y).clone()
    for _ in range(4):
        assert torch.equal(m(x, t, y), first)


def test_jvp_twice_bit_identical():
    m = _unet(dict(in_channels=3, num_classes=3), 78, "fp16")
    g = torch.Generator().manual_seed(4)
    B = 8
    x = torch.randn(B, 3, 32, 32, generator=g).to(DEV)
    v = torch.randn(B, 3, 32, 32, generator=g).to(DEV)
    t = torch.full((B,), 0.4, device=DEV)
    y = torch.full((B,), 1, device=DEV)
    e0, d0 = m.forward_jvp(x, t, y, v)
    e0, d0 = e0.clone(), d0.clone()
    for _ in range(3):
        e, d = m.forward_jvp(x, t, y, v)
        assert torch.equal(e, e0) and torch.equal(d, d0)


def test_chain_twice_bit_identical():
    from composable_diffusion_models_b200.compose_scores import sample_composed_sde
    experts = [_unet(dict(in_channels=1), s, "fp16") for s in (301, 302)]
    g = torch.Generator().manual_seed(5)
    B, n = 130, 30
    x0 = torch.randn(B, 1, 28, 28, generator=g)
    a = sample_composed_sde(experts, [0.5, 0.5], B, (1, 28, 28), n, 1.0, device=DEV, x_init=x0, noise="kernel", seed=11)
    b = sample_composed_sde(experts, [0.5, 0.5], B, (1, 28, 28), n, 1.0, device=DEV, x_init=x0, noise="kernel", seed=11)
    assert torch.equal(a, b)


def test_guided_unet_twice_bit_identical():
    from composable_diffusion_models_b200.models import GuidedUNet
    m = GuidedUNet(precision="fp16")
    m.load_state_dict(E.synth_state_dict(E.guided_unet_spec(), 5), strict=True)
    m = m.to(DEV).eval()
    g = torch.Generator().manual_seed(6)
    B = 33
    x = torch.randn(B, 3, 32, 32, generator=g).to(DEV)
    t = torch.full((B,), 250.0, device=DEV)
    d = torch.randint(0, 11, (B,), generator=g).to(DEV)
    c = torch.randint(0, 4, (B,), generator=g).to(DEV)
    first = m(x, t, d, c).clone()
    for _ in range(3):
        assert torch.equal(m(x, t, d, c), first)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_step_and_forward_on_a_non_current_device():
    """a tensor on cuda:1 must be stepped on cuda:1 while cuda:0 is the current device."""
    from composable_diffusion_models_b200 import steps
    torch.cuda.set_device(0)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(4, 1, 28, 28, generator=g)
    e = torch.randn(4, 1, 28, 28, generator=g)
    z = torch.randn(4, 1, 28, 28, generator=g)
    want = steps.step_sde(x.to("cuda:0"), [e.to("cuda:0")], [1.0], 0.3, 0.7, 1e-3, 0.1, z=z.to("cuda:0")).cpu()
    got = steps.step_sde(x.to("cuda:1"), [e.to("cuda:1")], [1.0], 0.3, 0.7, 1e-3, 0.1, z=z.to("cuda:1"))
    assert got.device.index == 1 and torch.equal(got.cpu(), want)
    m = _unet(dict(in_channels=1), 9, "fp16")
    t = torch.full((4,), 0.5)
    a = m(x.to("cuda:0"), t.to("cuda:0")).cpu()
    m = m.to("cuda:1")          # the native handle follows the module to the new device
    b = m(x.to("cuda:1"), t.to("cuda:1")).cpu()
    assert torch.equal(a, b)


PAIR_OPTIONS = ((b"conv_pair", 2), (b"conv_pair64", 1), (b"stack_pair", 1))     # every paired instance on (defaults pair only where it wins)


@pytest.mark.parametrize("cin,S,B", [(1, 28, 301), (3, 64, 9), (1, 28, 1), (1, 28, 2)])
def test_cta_pair_instances_equal_single_cta(cin, S, B):
    """CTA pairs (tcgen05 cta_group::2: the PAIR instances of csrc/conv_tc2.cu and conv_tc3.cu) only change which SM computes
    which tile and which CTA feeds which half of the weight rows, never the K order of an accumulator, so the forward is
    bit-identical to the single-CTA instances -- including odd tile counts, where the peer CTA of the last pair runs a
    dropped duplicate."""
    from composable_diffusion_models_b200 import _lib
    lib = _lib.lib()
    nc = 3 if cin == 3 else None
    m = _unet(dict(in_channels=cin, num_classes=nc), 79, "fp16")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, cin, S, S, generator=g).to(DEV)
    t = (torch.rand(B, generator=g) * 0.9 + 0.05).to(DEV)
    y = torch.randint(0, 3, (B,), generator=g).to(DEV) if nc else None
    try:
        for name, _ in PAIR_OPTIONS:
            _lib.check(lib.cdm_set_option(name, 0))
        single = m(x, t, y).clone()
        for name, v in PAIR_OPTIONS:
            _lib.check(lib.cdm_set_option(name, v))
        pair = m(x, t, y).clone()
    finally:
        for name, _ in PAIR_OPTIONS:
            lib.cdm_set_option(name, -1)
    assert torch.isfinite(pair).all()
    assert torch.equal(single, pair)


def test_cta_pair_grouped_chain_equals_single_cta():
    """The same through the grouped K = 2 chain entry (pair instances of the grouped kernels, gridDim.y = expert)."""
    from composable_diffusion_models_b200 import _lib
    from composable_diffusion_models_b200.compose_scores import sample_composed_sde
    lib = _lib.lib()
    experts = [_unet(dict(in_channels=1), 80 + k, "fp16") for k in range(2)]
    g = torch.Generator().manual_seed(6)
    B, n = 37, 4
    x0 = torch.randn(B, 1, 28, 28, generator=g)
    noise = torch.randn(n, B, 1, 28, 28, generator=g)
    outs = []
    try:
        for on in (0, 1):
            for name, v in PAIR_OPTIONS:
                _lib.check(lib.cdm_set_option(name, v if on else 0))
            outs.append(sample_composed_sde(experts, [0.5, 0.5], B, (1, 28, 28), n, 1.0, device=DEV, x_init=x0, noise=noise).clone())
    finally:
        for name, _ in PAIR_OPTIONS:
            lib.cdm_set_option(name, -1)
    assert torch.equal(outs[0], outs[1])
