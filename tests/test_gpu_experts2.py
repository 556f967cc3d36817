"""GPU: the BatchNorm score UNet (row a8, fp32 path) and the cross-attention GuidedUNet (row a7, fp32 path <= 1e-5 and
fp16 tensor-core path <= 2e-3) on libcdm_b200 vs the reference's golden outputs, the oracle at other sizes, and the full
SuperDiff / CFG samplers with NATIVE experts."""
import types

import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import experts as E
from oracle import samplers as OS
from oracle import schedule as S

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-5


def _score(seed, precision="fp32"):
    from composable_diffusion_models_b200.models import ColoredMNISTScoreModel
    m = ColoredMNISTScoreModel(precision=precision)
    sd = E.synth_state_dict(E.score_model_spec(), seed)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval(), sd


TOL_F16 = 2e-3


def _guided(seed, precision="fp32"):
    from composable_diffusion_models_b200.models import GuidedUNet
    m = GuidedUNet(precision=precision)
    sd = E.synth_state_dict(E.guided_unet_spec(), seed)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval(), sd


@pytest.mark.parametrize("precision,tol", [("fp32", TOL), ("fp16", TOL_F16)])
def test_score_model_vs_reference_golden(precision, tol):
    """fp16: every 3x3, k4-s2 strided and k4-s2 transposed conv on tcgen05 (conv_x3.cu TERMS = 1); 32-channel tensors
    zero-padded to 64 channels."""
    g = load_golden("score_model")
    m, _ = _score(g["seed"], precision)
    assert rel_l2(m(g["x"].to(DEV), g["t"].to(DEV)).cpu(), g["eps"]) < tol


@pytest.mark.parametrize("precision,tol", [("fp32", TOL), ("fp16", TOL_F16)])
@pytest.mark.parametrize("B,S", [(1, 32), (5, 16), (3, 64), (37, 32), (2, 8)])
def test_score_model_vs_oracle_sizes(B, S, precision, tol):
    m, sd = _score(700 + S, precision)
    g = torch.Generator().manual_seed(B + S)
    x = torch.randn(B, 3, S, S, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g).float()
    assert rel_l2(m(x.to(DEV), t.to(DEV)).cpu(), E.score_model_forward(sd, x, t)) < tol


def test_score_model_requires_eval_mode():
    m, _ = _score(1)
    m.train()
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 3, 32, 32, device=DEV), torch.zeros(1, device=DEV))


@pytest.mark.parametrize("precision,tol", [("fp32", TOL), ("fp16", TOL_F16)])
def test_guided_unet_vs_reference_golden(precision, tol):
    g = load_golden("guided_unet")
    m, _ = _guided(g["seed"], precision)
    got = m(g["x"].to(DEV), g["t"].to(DEV), g["digits"].to(DEV), g["colors"].to(DEV)).cpu()
    assert rel_l2(got, g["eps"]) < tol


@pytest.mark.parametrize("precision,tol", [("fp32", TOL), ("fp16", TOL_F16)])
@pytest.mark.parametrize("B,S", [(1, 32), (6, 16), (37, 32)])
def test_guided_unet_vs_oracle_sizes(B, S, precision, tol):
    m, sd = _guided(800 + S, precision)
    g = torch.Generator().manual_seed(B * S)
    x = torch.randn(B, 3, S, S, generator=g)
    t = torch.randint(0, 500, (B,), generator=g)
    d = torch.randint(0, 11, (B,), generator=g)
    c = torch.randint(0, 4, (B,), generator=g)
    want = E.guided_unet_forward(sd, x, t, d, c)
    assert rel_l2(m(x.to(DEV), t.to(DEV), d.to(DEV), c.to(DEV)).cpu(), want) < tol


@pytest.mark.parametrize("op", ["or", "and", "avg"])
def test_superdiff_sampler_native_experts_vs_reference(op):
    from composable_diffusion_models_b200.diffusion import SuperDiffSampler
    from composable_diffusion_models_b200.schedule import VPSDE
    g = load_golden(f"sampler_superdiff_{op}")
    m1, _ = _score(g["seed1"])
    m2, _ = _score(g["seed2"])
    sampler = SuperDiffSampler(VPSDE(num_timesteps=g["T"], device=DEV))
    out = sampler.sample(m1, m2, 2, (3, 32, 32), DEV, operation=op.upper(), temp=g["temp"], bias=0.0, x_init=g["x_init"],
                         noise=g["noise"])
    assert rel_l2(out.cpu(), g["out"]) < TOL


def test_superdiff_k4_vs_oracle():
    """K = 4 experts (config 4 of BASELINE.json): OR-softmax over four running log-densities."""
    from composable_diffusion_models_b200.diffusion import SuperDiffSampler
    from composable_diffusion_models_b200.schedule import VPSDE
    seeds = (11, 12, 13, 14)
    ms, sds = zip(*[_score(s) for s in seeds])
    T, B = 10, 3
    g = torch.Generator().manual_seed(21)
    x0 = torch.randn(B, 3, 32, 32, generator=g)
    noise = torch.randn(T - 1, B, 3, 32, 32, generator=g)
    sde = S.VPSDETables(num_timesteps=T)
    want, want_q = OS.sample_superdiff(sde, [lambda x, t, sd=sd: E.score_model_forward(sd, x, t) for sd in sds], x0, noise, "OR", 1.0, 0.0)
    sampler = SuperDiffSampler(VPSDE(num_timesteps=T, device=DEV))
    out, logq = sampler.sample(None, None, B, (3, 32, 32), DEV, operation="OR", models=list(ms), x_init=x0, noise=noise,
                               return_log_q=True)
    assert rel_l2(out.cpu(), want) < TOL
    assert rel_l2(logq.cpu(), want_q) < 1e-4


def test_cfg_sampler_native_guided_unet_vs_reference():
    from composable_diffusion_models_b200.compositional_diffusion_with_cross_attention import sample_composed
    g = load_golden("sampler_cfg_x0")
    m, _ = _guided(g["seed"])
    cfg = types.SimpleNamespace(DEVICE=DEV, IMG_SIZE=32, TIMESTEPS=g["timesteps"], GUIDANCE_STRENGTH_SHAPE=7.5,
                                GUIDANCE_STRENGTH_COLOR=7.5)
    out = sample_composed(cfg, m, g["digit"], g["color"], x_init=g["x_init"])
    assert rel_l2(out.cpu(), g["out"]) < TOL
    # batched chains are independent: sample 0 of a batch of 3 equals the batch-1 run
    x3 = torch.cat([g["x_init"], torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(1))])
    out3 = sample_composed(cfg, m, g["digit"], g["color"], batch_size=3, x_init=x3)
    assert rel_l2(out3[:1].cpu(), g["out"]) < TOL


# ---- the general fp16 tensor-core convolution (conv_x3.cu TERMS = 1): every geometry the score / SimpleUnet graphs use ----
def _conv_t16(x1, w, bias, kind, x2=None, relu=False):
    import ctypes
    from composable_diffusion_models_b200 import _lib
    lib = _lib.lib()
    B, C1, H, W = x1.shape
    C2 = x2.shape[1] if x2 is not None else 0
    Cout = w.shape[1] if kind == 2 else w.shape[0]
    Ho, Wo = (H // 2, W // 2) if kind == 1 else ((2 * H, 2 * W) if kind == 2 else (H, W))
    out = torch.empty(B, Cout, Ho, Wo, device=DEV)
    wh = w.float().contiguous().cpu()
    x1d = x1.to(DEV).contiguous()
    x2d = x2.to(DEV).contiguous() if x2 is not None else None
    bd = bias.to(DEV).contiguous() if bias is not None else None
    _lib.check(lib.cdm_debug_conv_t16(_lib.ptr(x1d), _lib.ptr(x2d), ctypes.c_void_p(wh.data_ptr()), _lib.ptr(bd), _lib.ptr(out), B, C1, C2,
                                      Cout, H, W, kind, 1 if relu else 0, _lib.stream_of(out)))
    return out.cpu()


T16_CASES = [
    # kind, B, C1, C2, Cout, S
    (0, 3, 64, 0, 64, 32), (0, 2, 32, 0, 64, 32), (0, 2, 32, 32, 32, 32), (0, 5, 128, 128, 128, 8), (0, 2, 64, 64, 64, 16),
    (0, 1, 256, 0, 256, 8), (0, 2, 1024, 0, 512, 4), (0, 2, 64, 0, 128, 64),
    (1, 3, 64, 0, 64, 32), (1, 2, 128, 0, 128, 16), (1, 5, 256, 0, 256, 8), (1, 2, 1024, 0, 1024, 8), (1, 1, 128, 0, 128, 64),
    (2, 3, 256, 0, 128, 4), (2, 2, 128, 0, 64, 8), (2, 2, 64, 0, 32, 16), (2, 37, 64, 0, 64, 8), (2, 1, 512, 0, 512, 4),
    (3, 2, 128, 0, 64, 16),
]


@pytest.mark.parametrize("case", T16_CASES)
def test_conv_t16_layer(case):
    import torch.nn.functional as F
    kind, B, C1, C2, Cout, S = case
    g = torch.Generator().manual_seed(sum(case))
    x1 = torch.randn(B, C1, S, S, generator=g).half().float()
    x2 = torch.randn(B, C2, S, S, generator=g).half().float() if C2 else None
    cin = C1 + C2
    k = {0: 3, 1: 4, 2: 4, 3: 1}[kind]
    if kind == 2:
        w = (torch.randn(cin, Cout, k, k, generator=g) / (k * cin ** 0.5)).half().float()
    else:
        w = (torch.randn(Cout, cin, k, k, generator=g) / (k * cin ** 0.5)).half().float()
    bias = torch.randn(Cout, generator=g)
    xin = torch.cat([x1, x2], dim=1) if C2 else x1
    if kind == 0:
        want = F.conv2d(xin.double(), w.double(), bias.double(), padding=1)
    elif kind == 1:
        want = F.conv2d(xin.double(), w.double(), bias.double(), stride=2, padding=1)
    elif kind == 2:
        want = F.conv_transpose2d(xin.double(), w.double(), bias.double(), stride=2, padding=1)
    else:
        want = F.conv2d(xin.double(), w.double(), bias.double())
    got = _conv_t16(x1, w, bias, kind, x2)
    assert rel_l2(got, want) < 5e-4          # fp16 output rounding (operands are exactly representable)
    got = _conv_t16(x1, w, bias, kind, x2, relu=True)
    assert rel_l2(got, want.clamp_min(0)) < 5e-4
