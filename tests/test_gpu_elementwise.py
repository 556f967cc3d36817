"""GPU: the two elementwise producers of the UNet graphs in isolation, through the C-ABI hooks cdm_debug_maxpool / cdm_debug_upcat:
2x2 max pool (reference: self.pool = nn.MaxPool2d(2), mnist/models/unet_small.py:62,80,82) and bilinear x2 upsample with
align_corners=True + channel concat (self.unpool + torch.cat, :70,84-85,88-89), each with the GroupNorm {sum, sumsq} it
accumulates for its consumer -- including the virtual-concat mode (only the upsampled channels are written, the statistics still
cover the whole concat) and maps tall enough to split a sample over several CTAs.

fp32: max pool is exact; the upsample follows torch's expression tree (a few fp32 ulps).  fp16: inputs are rounded to fp16 on the
way in, so the references are computed from the rounded inputs and only the output rounding (2^-11) remains.  Statistics are
fixed-point sums of a fixed-order float tree: they must match the sums of the kernel's own output."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _maxpool(x, precision, want_in=True):
    from composable_diffusion_models_b200 import _lib
    lib = _lib.lib()
    B, C, H, W = x.shape
    out = torch.empty(B, C, H // 2, W // 2, device=DEV)
    st = torch.zeros(B, 8, 2, device=DEV)
    st_in = torch.zeros(B, 8, 2, device=DEV) if want_in else None
    xd = x.to(DEV).contiguous()
    _lib.check(lib.cdm_debug_maxpool(_lib.ptr(xd), _lib.ptr(out), _lib.ptr(st), _lib.ptr(st_in), B, C, H, W,
                                     _lib.precision_code(precision), _lib.stream_of(out)))
    return out.cpu(), st.cpu(), (st_in.cpu() if want_in else None)


def _upcat(low, skip, precision, virt):
    from composable_diffusion_models_b200 import _lib
    lib = _lib.lib()
    B, Ca, h, w = low.shape
    Cs = skip.shape[1]
    out = torch.empty(B, Ca if virt else Ca + Cs, 2 * h, 2 * w, device=DEV)
    st = torch.zeros(B, 8, 2, device=DEV)
    ld, sd = low.to(DEV).contiguous(), skip.to(DEV).contiguous()
    _lib.check(lib.cdm_debug_upcat(_lib.ptr(ld), _lib.ptr(sd), _lib.ptr(out), _lib.ptr(st), B, Ca, Cs, h, w,
                                   _lib.precision_code(precision), 1 if virt else 0, _lib.stream_of(out)))
    return out.cpu(), st.cpu()


def _group_sums(t):
    g = t.double().view(t.shape[0], 8, -1)
    return torch.stack([g.sum(-1), (g * g).sum(-1)], dim=-1).float()


POOL_CASES = [(5, 64, 28, 28), (3, 128, 14, 14), (2, 64, 64, 64), (300, 64, 4, 6), (1, 256, 2, 2), (2, 384, 8, 8)]


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
@pytest.mark.parametrize("case", POOL_CASES)
def test_maxpool_and_both_statistics(case, precision):
    B, C, H, W = case
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn(B, C, H, W, generator=g) * 2 + 0.3
    if precision == "fp16":
        x = x.half().float()
    out, st, st_in = _maxpool(x, precision)
    assert torch.equal(out, F.max_pool2d(x, 2))            # a max of representable values is exact in either precision
    assert rel_l2(st, _group_sums(out)) < 1e-5
    assert rel_l2(st_in, _group_sums(x)) < 1e-5


UPCAT_CASES = [
    # B, Ca, Cs, h, w, virtual
    (4, 256, 128, 7, 7, True),        # MNIST up1: bottleneck -> 14x14
    (3, 128, 64, 14, 14, True),       # MNIST up2 -> 28x28
    (2, 128, 64, 32, 32, True),       # shapes up2 -> 64x64
    (4, 256, 128, 7, 7, False),
    (3, 128, 64, 14, 14, False),
    (2, 256, 128, 16, 16, False),
    (300, 128, 64, 2, 3, False),      # more samples than SMs, tiny non-square maps
    (1, 64, 64, 1, 1, False),         # a single low-resolution pixel: scale 0
]


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
@pytest.mark.parametrize("case", UPCAT_CASES)
def test_upsample_concat_and_statistics(case, precision):
    B, Ca, Cs, h, w, virt = case
    g = torch.Generator().manual_seed(Ca + h)
    low = torch.randn(B, Ca, h, w, generator=g) * 1.5 - 0.2
    skip = torch.randn(B, Cs, 2 * h, 2 * w, generator=g) + 0.5
    if precision == "fp16":
        low, skip = low.half().float(), skip.half().float()
    up = F.interpolate(low, scale_factor=2, mode="bilinear", align_corners=True)
    want = torch.cat([up, skip], 1)
    out, st = _upcat(low, skip, precision, virt)
    tol = 2e-6 if precision == "fp32" else 2.5e-4
    assert rel_l2(out, up if virt else want) < tol
    # statistics of the WHOLE concat in either mode: the upsampled part as stored + the skip tensor
    whole = torch.cat([out[:, :Ca], skip], 1)
    assert rel_l2(st, _group_sums(whole)) < 2e-5


def test_elementwise_hooks_reject_bad_shapes():
    with pytest.raises((NotImplementedError, ValueError)):
        _maxpool(torch.randn(1, 64, 7, 8), "fp16")           # odd height
    with pytest.raises((NotImplementedError, ValueError)):
        _upcat(torch.randn(1, 100, 4, 4), torch.randn(1, 64, 8, 8), "fp16", True)      # no virtual concat for 100 channels
