"""GPU: the 62 M-parameter SimpleUnet (SURVEY.md section 8(f) row 4, fp32 path) on libcdm_b200 vs the reference-generated
golden, the oracle at another size / batch, and as the expert of the classifier-free-guidance SuperDiff sampler."""
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import experts as E

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-5


def _make(precision):
    from composable_diffusion_models_b200.models import SimpleUnet
    g = load_golden("simple_unet")
    m = SimpleUnet(g["num_classes"], precision=precision)
    sd = E.synth_state_dict(E.simple_unet_spec(g["num_classes"]), g["seed"])
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval(), sd


@pytest.fixture(scope="module")
def model():
    return _make("fp32")


@pytest.fixture(scope="module")
def model16():
    return _make("fp16")


def test_fp16_tensor_core_forward(model16):
    """The tcgen05 graph (3x3 / k4-s2 strided / k4-s2 transposed convs with bias + ReLU + GroupNorm statistics in their
    epilogues) against the reference-generated golden and the oracle at other sizes; bound 2e-3 like the other fp16 experts."""
    m, sd = model16
    g = load_golden("simple_unet")
    got = m(g["x"].to(DEV), g["t"].to(DEV), g["y"].to(DEV))
    assert rel_l2(got.cpu(), g["out"]) < 2e-3
    for B, S in ((1, 16), (5, 48), (2, 64), (3, 32)):
        gen = torch.Generator().manual_seed(B * S)
        x = torch.randn(B, 3, S, S, generator=gen)
        t = torch.randint(0, 500, (B,), generator=gen)
        y = torch.randint(0, 4, (B,), generator=gen)
        got = m(x.to(DEV), t.to(DEV), y.to(DEV))
        assert rel_l2(got.cpu(), E.simple_unet_forward(sd, x, t, y)) < 2e-3, (B, S)
        assert torch.equal(got, m(x.to(DEV), t.to(DEV), y.to(DEV)))          # run-to-run bit-identical


def test_forward_vs_reference_golden(model):
    m, _ = model
    g = load_golden("simple_unet")
    got = m(g["x"].to(DEV), g["t"].to(DEV), g["y"].to(DEV))
    assert rel_l2(got.cpu(), g["out"]) < TOL


@pytest.mark.parametrize("B,S", [(1, 16), (5, 48), (2, 64)])
def test_forward_vs_oracle(model, B, S):
    m, sd = model
    g = torch.Generator().manual_seed(B * S)
    x = torch.randn(B, 3, S, S, generator=g)
    t = torch.randint(0, 500, (B,), generator=g)
    y = torch.randint(0, 4, (B,), generator=g)
    got = m(x.to(DEV), t.to(DEV), y.to(DEV))
    assert rel_l2(got.cpu(), E.simple_unet_forward(sd, x, t, y)) < TOL


def test_micro_batching_and_edge_cases(model, monkeypatch):
    m, sd = model
    g = torch.Generator().manual_seed(3)
    x = torch.randn(7, 3, 16, 16, generator=g)
    t = torch.randint(0, 500, (7,), generator=g)
    y = torch.randint(0, 4, (7,), generator=g)
    want = E.simple_unet_forward(sd, x, t, y)
    monkeypatch.setenv("CDM_SIMPLE_MICROBATCH", "3")           # 3 + 3 + 1 samples through the same workspace
    assert rel_l2(m(x.to(DEV), t.to(DEV), y.to(DEV)).cpu(), want) < TOL
    monkeypatch.delenv("CDM_SIMPLE_MICROBATCH")
    assert m(torch.empty(0, 3, 16, 16, device=DEV), torch.empty(0, device=DEV), torch.empty(0, dtype=torch.long, device=DEV)).shape[0] == 0
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 3, 24, 24, device=DEV), torch.zeros(1, device=DEV), torch.zeros(1, dtype=torch.long, device=DEV))
    with pytest.raises(ValueError):
        m(torch.zeros(1, 1, 16, 16, device=DEV), torch.zeros(1, device=DEV), torch.zeros(1, dtype=torch.long, device=DEV))


def test_as_superdiff_expert(model):
    """The use the reference makes of it: conditional + null-token forwards inside sample_superdiff
    (src/..._shape_and_color_6_1.py:331-430); here through this repo's K-expert linear-solve sampler vs the oracle."""
    from composable_diffusion_models_b200 import superdiff_linear_solve as SL
    from oracle import samplers as OS
    m, sd = model
    T, S = 4, 16
    g = torch.Generator().manual_seed(11)
    x0 = torch.randn(2, 3, S, S, generator=g)
    dw = torch.randn(T, 2, 3, S, S, generator=g)
    zs = torch.randn(T - 1, 2, 3, S, S, generator=g)
    labels = (0, 2)
    import types
    oracle = [lambda img, t, k=k: E.simple_unet_forward(sd, img, t, torch.full((img.shape[0],), labels[k])) for k in range(2)]
    want, _ = OS.sample_superdiff_6_1(T, oracle, x0, dw, zs, "AND", 1.0, 0.0)
    cfg = types.SimpleNamespace(DEVICE=DEV, TIMESTEPS=T, IMG_SIZE=S)
    got = SL.sample_superdiff(m, m, labels[0], labels[1], mode="AND", T=1.0, l=0.0, batch_size=2, x_init=x0, dw=dw, noise=zs,
                              config=cfg)
    assert rel_l2(got.cpu(), want) < 1e-4
