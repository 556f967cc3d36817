"""GPU: the init conv of the fp16 expert graphs in isolation (reference: UNet.init_conv = Conv2d(Cin, 64, 3, padding=1),
mnist/models/unet_small.py:57,78, shapes/models/unet_small.py:74,106), through the C-ABI hook cdm_debug_init_conv: the tcgen05
kernel (pixel-major hi/lo image buffer, two taps per K = 16 MMA, no im2col) and the CUDA-core kernel against F.conv2d.

The input image and the weights keep fp32 precision in both kernels (hi + lo fp16 splits on the tensor core), so the only
error is the fp16 rounding of the stored output (2^-11 per value, ~1.6e-4 rel-L2); GroupNorm statistics are taken from the
fp32 accumulators (tensor-core kernel) or the rounded values (CUDA-core kernel) and must match the output's own sums."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _init_conv(x, w, bias, tc, want_stats=True):
    from composable_diffusion_models_b200 import _lib
    lib = _lib.lib()
    B, Cin, H, W = x.shape
    out = torch.empty(B, 64, H, W, device=DEV)
    stats = torch.zeros(B, 8, 2, device=DEV) if want_stats else None
    xd, wd = x.to(DEV).contiguous(), w.to(DEV).contiguous()
    bd = bias.to(DEV).contiguous() if bias is not None else None
    _lib.check(lib.cdm_debug_init_conv(_lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(out), _lib.ptr(stats), B, Cin, H, W,
                                       1 if tc else 0, _lib.stream_of(out)))
    return out.cpu(), (stats.cpu() if want_stats else None)


CASES = [
    # B, Cin, H, W
    (5, 1, 28, 28),      # MNIST (C2): 7 tiles per sample, the last one ragged
    (3, 3, 64, 64),      # shapes (C3 / C4): 33 tiles per sample
    (4, 3, 28, 28),      # colored MNIST (C5)
    (2, 3, 32, 32),      # score-model size
    (300, 1, 28, 28),    # more samples than SMs: the double-buffered image ring wraps, CTAs take 2-3 samples
    (1, 2, 9, 13),       # odd, non-square, a single tile
    (2, 3, 5, 40),       # wide and short
    (1, 1, 1, 1),        # one pixel
]


@pytest.mark.parametrize("tc", [True, False])
@pytest.mark.parametrize("case", CASES)
def test_init_conv_vs_conv2d(case, tc):
    B, Cin, H, W = case
    g = torch.Generator().manual_seed(B * 1000 + Cin * 100 + H)
    x = torch.randn(B, Cin, H, W, generator=g) * 1.5
    w = torch.randn(64, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)
    bias = torch.randn(64, generator=g)
    want = F.conv2d(x.double(), w.double(), bias.double(), padding=1)
    got, stats = _init_conv(x, w, bias, tc)
    # fp16 output rounding only: 2^-11 / sqrt(3) per value
    assert rel_l2(got, want) < 2.5e-4, rel_l2(got, want)
    # against the fp16-rounded exact result the kernels may differ by one fp16 ulp on rare ties only
    assert (got - want.float().half().float()).abs().max() <= want.abs().max().item() * 2 ** -10
    gv = (want if tc else got.double()).view(B, 8, -1)        # tensor core: sums of the fp32 accumulators; CUDA cores: of the stored values
    assert rel_l2(stats[:, :, 0], gv.sum(-1).float()) < 2e-4 and rel_l2(stats[:, :, 1], (gv * gv).sum(-1).float()) < 2e-4


def test_init_conv_tc_no_bias_and_repeat_bit_identical():
    g = torch.Generator().manual_seed(5)
    x = torch.randn(7, 3, 64, 64, generator=g)
    w = torch.randn(64, 3, 3, 3, generator=g) / 5
    a, sa = _init_conv(x, w, None, True)
    b, sb = _init_conv(x, w, None, True)
    assert torch.equal(a, b) and torch.equal(sa, sb)
    assert rel_l2(a, F.conv2d(x.double(), w.double(), padding=1)) < 2.5e-4


def test_init_conv_tc_keeps_fp32_inputs():
    """Values that fp16 cannot hold (a large offset plus a small signal) survive the hi/lo split of image and weights."""
    g = torch.Generator().manual_seed(6)
    x = 3.0 + 1e-3 * torch.randn(2, 3, 28, 28, generator=g)
    w = torch.randn(64, 3, 3, 3, generator=g) / 5
    w = w - w.mean(dim=(1, 2, 3), keepdim=True)            # zero-sum filters: the offset cancels, the small signal remains
    want = F.conv2d(x.double(), w.double(), padding=1)
    got, _ = _init_conv(x, w, None, True, want_stats=False)
    inner = (slice(None), slice(None), slice(1, -1), slice(1, -1))      # away from the zero padding the output is O(1e-3)
    assert rel_l2(got[inner], want[inner]) < 2e-3
    # a single-product fp16 conv (x rounded to fp16: ulp(3.0) = 2e-3) could not resolve this signal at all
    naive = F.conv2d(x.half().double(), w.half().double(), padding=1)
    assert rel_l2(naive[inner].float(), want[inner]) > 0.3


def test_init_conv_tc_unsupported_shapes_raise():
    x = torch.randn(1, 4, 8, 8)
    w = torch.randn(64, 4, 3, 3)
    with pytest.raises(NotImplementedError):
        _init_conv(x, w, None, True)
