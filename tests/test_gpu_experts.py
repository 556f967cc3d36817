"""GPU: expert denoisers on libcdm_b200 vs the oracle and the reference's golden outputs.

fp32 path: <= 1e-5 rel-L2 (north_star's fp32 bound).  fp16 tcgen05 path: single forward <= 2e-3 rel-L2
against fp32 (fp16 operands and activation storage, fp32 accumulation; measured 0.5e-3 .. 1.25e-3, i.e. ~25 roundings
of 2^-12 each; bf16 measured 5e-3; the end-to-end bound is exercised in test_gpu_samplers)."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_l2
from oracle import experts as E

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL_FP32 = 1e-5
TOL_F16 = 2e-3


def _debug_conv(x, w, bias, res=None, wres=None, identity=None, taps=9, precision="fp32", want_stats=False):
    from composable_diffusion_models_b200 import _lib
    lib = _lib.lib()
    B, Cin, H, W = x.shape
    Cout = w.shape[0]
    out = torch.empty(B, Cout, H, W, device=DEV)
    stats = torch.zeros(B, 8, 2, device=DEV) if want_stats else None
    wh = w.float().contiguous().cpu()
    wrh = wres.float().contiguous().cpu() if wres is not None else None
    xd, bd = x.to(DEV).contiguous(), bias.to(DEV).contiguous()
    rd = res.to(DEV).contiguous() if res is not None else None
    idd = identity.to(DEV).contiguous() if identity is not None else None
    _lib.check(lib.cdm_debug_conv(_lib.ptr(xd), ctypes.c_void_p(wh.data_ptr()), _lib.ptr(bd), bias.shape[0] if bias.dim() == 2 else 1,
                                  _lib.ptr(rd), ctypes.c_void_p(wrh.data_ptr()) if wrh is not None else None, _lib.ptr(idd),
                                  _lib.ptr(out), _lib.ptr(stats), B, Cin, res.shape[1] if res is not None else 0, Cout, H, W, taps,
                                  {"fp16_halo": 2, "fp16_stack": 3}.get(precision) or _lib.precision_code(precision), _lib.stream_of(out)))
    return out.cpu(), (stats.cpu() if want_stats else None)


CONV_CASES = [
    # B, Cin, Cout, S, Cres, identity   (the layer shapes of the mnist / shapes UNets)
    (3, 64, 64, 28, 0, True),
    (5, 64, 128, 14, 0, False),
    (3, 128, 128, 14, 64, False),
    (5, 128, 256, 7, 0, False),
    (3, 256, 256, 7, 128, False),
    (2, 384, 128, 14, 0, False),
    (2, 64, 64, 28, 192, False),
    (2, 64, 64, 64, 0, True),
    (2, 128, 128, 32, 64, False),
    (3, 256, 256, 16, 128, False),
    (1, 64, 64, 8, 0, False),
    (33, 128, 128, 14, 0, False),     # more samples than one 2x2x32 tile holds
    (2, 192, 64, 28, 0, False),       # up2.conv1: three 64-channel K chunks
    (3, 64, 64, 32, 0, True),
    (1, 128, 128, 16, 0, False),
    (5, 64, 128, 32, 0, False),
    # stacked-tap kernel coverage: many tiles per CTA, resident and streamed weight tiles, ragged last tile, 128-row strips
    (40, 64, 64, 28, 64, False),
    (30, 128, 64, 28, 0, True),
    (2, 64, 64, 30, 0, False),
    (4, 64, 64, 14, 0, False),
    (3, 64, 64, 20, 256, False),
    (3, 64, 64, 19, 0, True),         # odd size: ragged last 4-row tile, 13 zero-fill columns
    # two samples per tile (halo kernel scheme C): odd / even / single batches, smaller-than-7 maps, identity, long K
    (1, 128, 256, 7, 0, False),
    (8, 256, 256, 7, 128, False),
    (5, 256, 256, 7, 0, True),
    (3, 64, 256, 5, 64, False),
    (4, 128, 256, 4, 0, False),
]


def _stack_ok(Cin, Cout, S, Cres):
    return Cout == 64 and not (S % 8 == 0 and S % 16 == 0) and 20 < S + 2 <= 32 and S * S >= 196


@pytest.mark.parametrize("precision", ["fp32", "f16x3", "fp16", "fp16_halo", "fp16_stack"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_layer(case, precision):
    B, Cin, Cout, S, Cres, ident = case
    if precision == "fp16_halo" and S < 14 and not (S <= 7 and Cout == 256):
        pytest.skip("halo-tile kernel handles maps >= 14x14 and (two samples per tile) <= 7x7 with 256 output channels")
    if precision == "fp16_stack" and not _stack_ok(Cin, Cout, S, Cres):
        pytest.skip("stacked-tap kernel: Cout = 64 full-width strips only")
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = torch.randn(B, Cin, S, S, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)
    bias = torch.randn(B, Cout, generator=g)
    res = torch.randn(B, Cres, S, S, generator=g) if Cres else None
    wres = torch.randn(Cout, Cres, generator=g) / Cres ** 0.5 if Cres else None
    idn = torch.randn(B, Cout, S, S, generator=g) if ident else None
    if precision not in ("fp32", "f16x3"):   # compare against the same fp16-rounded operands, so only accumulation order differs
        x, w = x.half().float(), w.half().float()
        if res is not None:
            res, wres = res.half().float(), wres.half().float()
        if idn is not None:
            idn = idn.half().float()
    want = F.conv2d(x.double(), w.double(), padding=1) + bias[:, :, None, None].double()
    if res is not None:
        want = want + F.conv2d(res.double(), wres[:, :, None, None].double())
    if idn is not None:
        want = want + idn.double()
    got, stats = _debug_conv(x, w, bias, res, wres, idn, precision=precision, want_stats=True)
    # fp32: fp32 accumulation error only.  f16x3 (three-term split-fp16 on tcgen05): exact products, but the tensor core
    # accumulates with truncation -- measured 7.7e-10 * K relative (3.9e-7 at K = 576 ... 2.7e-6 at K = 3456; with one
    # shared accumulator for all three terms it was 3x that).  fp16: 2^-11 output rounding.
    tol = {"fp32": 2e-6, "f16x3": 4e-6}.get(precision, 5e-4)
    assert rel_l2(got, want) < tol
    gv = got.view(B, 8, -1)
    assert rel_l2(stats[:, :, 0], gv.sum(-1)) < 1e-4 and rel_l2(stats[:, :, 1], (gv * gv).sum(-1)) < 1e-4


@pytest.mark.parametrize("precision", ["fp16_halo", "fp16_stack"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_layer_cta_pairs_bit_identical(case, precision):
    """Every layer shape on the CTA-pair instances (cta_group::2) == the single-CTA instances, outputs and statistics."""
    from composable_diffusion_models_b200 import _lib
    B, Cin, Cout, S, Cres, ident = case
    if precision == "fp16_halo" and (S < 14 or Cout == 256):
        pytest.skip("no paired instance for this shape")
    if precision == "fp16_stack" and not _stack_ok(Cin, Cout, S, Cres):
        pytest.skip("stacked-tap kernel: Cout = 64 full-width strips only")
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = torch.randn(B, Cin, S, S, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)
    bias = torch.randn(B, Cout, generator=g)
    res = torch.randn(B, Cres, S, S, generator=g) if Cres else None
    wres = torch.randn(Cout, Cres, generator=g) / Cres ** 0.5 if Cres else None
    idn = torch.randn(B, Cout, S, S, generator=g) if ident else None
    lib = _lib.lib()
    opts = ((b"conv_pair", 2), (b"conv_pair64", 1), (b"stack_pair", 1))
    outs = []
    try:
        for on in (0, 1):
            for name, v in opts:
                _lib.check(lib.cdm_set_option(name, v if on else 0))
            outs.append(_debug_conv(x, w, bias, res, wres, idn, precision=precision, want_stats=True))
    finally:
        for name, _ in opts:
            lib.cdm_set_option(name, -1)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("precision", ["fp32", "f16x3", "fp16"])
def test_conv_1x1(precision):
    g = torch.Generator().manual_seed(11)
    x = torch.randn(4, 128, 16, 16, generator=g).half().float()
    w = (torch.randn(64, 128, 1, 1, generator=g) / 11).half().float()
    bias = torch.randn(1, 64, generator=g)
    got, _ = _debug_conv(x, w, bias, taps=1, precision=precision)
    assert rel_l2(got, F.conv2d(x, w) + bias[:, :, None, None]) < (2e-6 if precision in ("fp32", "f16x3") else 4e-3)


def _native_unet(kw, seed, precision):
    from composable_diffusion_models_b200.models import UNet
    m = UNet(**kw, precision=precision)
    sd = E.synth_state_dict(E.unet_small_spec(kw.get("in_channels", 1), num_classes=kw.get("num_classes")), seed)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval(), sd


@pytest.mark.parametrize("precision,tol", [("fp32", TOL_FP32), ("f16x3", TOL_FP32), ("fp16", TOL_F16)])
def test_unet_mnist_vs_reference_golden(precision, tol):
    g = load_golden("unet_mnist")
    m, sd = _native_unet(dict(in_channels=1), g["seed"], precision)
    got = m(g["x"].to(DEV), g["t"].to(DEV)).cpu()
    assert rel_l2(got, g["eps"]) < tol
    if precision == "fp32":     # layer-by-layer against the oracle's intermediates
        _, mid = E.unet_small_forward(sd, g["x"], g["t"], return_intermediates=True)
        for name in ("x0", "d1", "d2", "b1", "u1", "u2"):
            assert rel_l2(m.debug_read(name, 3, 28).cpu(), mid[name]) < tol, name


@pytest.mark.parametrize("precision,tol", [("fp32", TOL_FP32), ("f16x3", TOL_FP32), ("fp16", TOL_F16)])
def test_unet_shapes_conditional_vs_reference_golden(precision, tol):
    g = load_golden("unet_shapes")
    ms, _ = _native_unet(dict(in_channels=1, num_classes=3), g["seed_shape"], precision)
    mc, _ = _native_unet(dict(in_channels=3, num_classes=3), g["seed_color"], precision)
    y = g["y"].to(DEV)
    assert rel_l2(ms(g["x_shape"].to(DEV), g["t"].to(DEV), y).cpu(), g["eps_shape"]) < tol
    assert rel_l2(mc(g["x_color"].to(DEV), g["t"].to(DEV), y).cpu(), g["eps_color"]) < tol
    with pytest.raises(ValueError):
        ms(g["x_shape"].to(DEV), g["t"].to(DEV))


@pytest.mark.parametrize("precision,tol", [("fp32", TOL_FP32), ("f16x3", TOL_FP32), ("fp16", TOL_F16)])
@pytest.mark.parametrize("B,S,cin", [(1, 28, 1), (37, 28, 1), (2, 64, 3), (9, 16, 3)])
def test_unet_vs_oracle_sizes(B, S, cin, precision, tol):
    """Ragged batch sizes (not multiples of any tile) and the 64x64 three-channel expert."""
    nc = 3 if cin == 3 else None
    m, sd = _native_unet(dict(in_channels=cin, num_classes=nc), 500 + S, precision)
    g = torch.Generator().manual_seed(B * S)
    x = torch.randn(B, cin, S, S, generator=g)
    t = torch.rand(B, generator=g) * 0.98 + 0.01
    y = torch.randint(0, 3, (B,), generator=g) if nc else None
    want = E.unet_small_forward(sd, x, t, y)
    got = m(x.to(DEV), t.to(DEV), y.to(DEV) if nc else None).cpu()
    assert rel_l2(got, want) < tol


def test_unet_microbatching_is_invisible(monkeypatch):
    """A batch larger than the micro-batch must give the same numbers as sample-by-sample evaluation."""
    from composable_diffusion_models_b200 import _lib
    m, sd = _native_unet(dict(in_channels=1), 77, "fp32")
    g = torch.Generator().manual_seed(0)
    B = 600
    _lib.lib().cdm_set_microbatch(512)       # force the batch to be split (default micro-batch is 4096)
    x = torch.randn(B, 1, 28, 28, generator=g).to(DEV)
    t = torch.rand(B, generator=g).to(DEV)
    full = m(x, t)
    part = torch.cat([m(x[i:i + 200], t[i:i + 200]) for i in range(0, B, 200)])
    assert rel_l2(full.cpu(), part.cpu()) < 1e-6
    idx = torch.tensor([0, 511, 512, 599])
    want = E.unet_small_forward(sd, x[idx].cpu(), t[idx].cpu())
    assert rel_l2(full[idx].cpu(), want) < TOL_FP32
    _lib.lib().cdm_set_microbatch(0)


@pytest.mark.parametrize("cin,S", [(1, 28), (3, 64), (3, 32)])
def test_f16_kernel_variants_agree(cin, S):
    """The three fp16 conv configurations (shifted-box kernel; halo-tile kernel with a separate GroupNorm pass;
    halo-tile kernel with GroupNorm+SiLU fused into its prologue) must agree to fp16 rounding."""
    from composable_diffusion_models_b200 import _lib
    lib = _lib.lib()
    nc = 3 if cin == 3 else None
    m, sd = _native_unet(dict(in_channels=cin, num_classes=nc), 321, "fp16")
    g = torch.Generator().manual_seed(5)
    B = 5
    x = torch.randn(B, cin, S, S, generator=g).to(DEV)
    t = (torch.rand(B, generator=g) * 0.9 + 0.05).to(DEV)
    y = torch.randint(0, 3, (B,), generator=g).to(DEV) if nc else None
    outs = {}
    try:
        for name, halo, fuse, stack in (("box", 0, 0, 0), ("halo", 1, 0, 0), ("halo+gn", 1, 1, 0), ("stack+gn", 1, 1, 1)):
            lib.cdm_set_option(b"conv_halo", halo)
            lib.cdm_set_option(b"fuse_gn", fuse)
            lib.cdm_set_option(b"conv_stack", stack)
            outs[name] = m(x, t, y).cpu()
    finally:
        lib.cdm_set_option(b"conv_halo", -1)
        lib.cdm_set_option(b"fuse_gn", -1)
        lib.cdm_set_option(b"conv_stack", -1)
    want = E.unet_small_forward(sd, x.cpu(), t.cpu(), y.cpu() if nc else None)
    for name, o in outs.items():
        assert rel_l2(o, want) < TOL_F16, name
    assert rel_l2(outs["halo"], outs["box"]) < 2e-3
    assert rel_l2(outs["halo+gn"], outs["halo"]) < 2e-3
    assert rel_l2(outs["stack+gn"], outs["halo+gn"]) < 2e-3


def test_unet_reloads_after_parameter_update():
    m, _ = _native_unet(dict(in_channels=1), 5, "fp32")
    x = torch.randn(2, 1, 28, 28, device=DEV)
    t = torch.tensor([0.5, 0.5], device=DEV)
    a = m(x, t)
    sd2 = E.synth_state_dict(E.unet_small_spec(1), 6)
    m.load_state_dict(sd2)
    b = m(x, t)
    assert rel_l2(b.cpu(), E.unet_small_forward(sd2, x.cpu(), t.cpu())) < TOL_FP32
    assert rel_l2(a.cpu(), b.cpu()) > 1e-2


def test_mlp_vs_reference_golden():
    from composable_diffusion_models_b200.models import MLP
    g = load_golden("mlp_2d")
    m = MLP()
    m.load_state_dict(E.synth_state_dict(E.mlp_2d_spec(), g["seed"]), strict=True)
    m = m.to(DEV)
    assert rel_l2(m(g["t"].to(DEV), g["x"].to(DEV)).cpu(), g["eps"]) < TOL_FP32
    gg = torch.Generator().manual_seed(1)
    x, t = torch.randn(1000, 2, generator=gg), torch.rand(1000, generator=gg)
    want = E.mlp_2d_forward(E.synth_state_dict(E.mlp_2d_spec(), g["seed"]), t, x)
    assert rel_l2(m(t.to(DEV), x.to(DEV)).cpu(), want) < TOL_FP32
