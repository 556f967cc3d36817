"""The small GroupNorm ResBlock UNet expert, B200-native.

Same constructor, ``forward(x, t[, y])`` and ``state_dict()`` keys as the reference's
``mnist/models/unet_small.py:47-92`` (unconditional) and ``shapes/models/unet_small.py:53-120``
(``num_classes`` adds ``label_emb``; a missing ``y`` raises ``ValueError`` as at :99-101).  The module
only *holds* the parameters (so ``load_state_dict(strict=True)`` and the reference's
``load_checkpoint`` work unchanged); the forward pass is ``cdm_unet_forward`` in libcdm_b200.so.

``precision``: "fp16" (tcgen05/TMA implicit-GEMM convolutions, default) or "fp32" (CUDA-core path that
tracks the fp32 reference to ~1e-6).
"""
import ctypes as C
import math
import os

import torch
import torch.nn as nn

from .. import _lib
from . import _native


def _res_block_params(cin, cout, tdim):
    # registration order follows the reference's ResBlock so default initialisation consumes the
    # global RNG identically (block1 -> time_mlp -> block2 -> res_conv)
    blk = nn.Module()
    blk.block1 = nn.ModuleDict({"0": nn.GroupNorm(8, cin), "2": nn.Conv2d(cin, cout, kernel_size=3, padding=1)})
    blk.time_mlp = nn.ModuleDict({"1": nn.Linear(tdim, cout)})
    blk.block2 = nn.ModuleDict({"0": nn.GroupNorm(8, cout), "3": nn.Conv2d(cout, cout, kernel_size=3, padding=1)})
    if cin != cout:
        blk.res_conv = nn.Conv2d(cin, cout, 1)
    return blk


class UNet(_native.NativeModule):
    _abi = "cdm_unet"
    _train_note = "the reference ResBlock's Dropout(0.1) is not applied"

    def __init__(self, in_channels=1, base_dim=64, time_emb_dim=256, num_classes=None, precision=None):
        super().__init__()
        self.in_channels, self.base_dim, self.time_emb_dim = in_channels, base_dim, time_emb_dim
        self.num_classes = num_classes
        self.precision = precision or os.environ.get("CDM_PRECISION", "fp16")
        d = base_dim
        self.time_mlp = nn.ModuleDict({"1": nn.Linear(d, time_emb_dim), "3": nn.Linear(time_emb_dim, time_emb_dim)})
        if num_classes is not None:
            self.label_emb = nn.Embedding(num_classes, time_emb_dim)
        self.init_conv = nn.Conv2d(in_channels, d, kernel_size=3, padding=1)
        self.down1 = _res_block_params(d, d, time_emb_dim)
        self.down2 = _res_block_params(d, 2 * d, time_emb_dim)
        self.bot1 = _res_block_params(2 * d, 4 * d, time_emb_dim)
        self.up1 = _res_block_params(6 * d, 2 * d, time_emb_dim)
        self.up2 = _res_block_params(3 * d, d, time_emb_dim)
        self.out_conv = nn.Conv2d(d, in_channels, kernel_size=1)

    # -- native handle ---------------------------------------------------------------------------
    def _create_native(self, lib, device_index):
        cfg = _lib.UNetConfig(self.in_channels, self.base_dim, self.time_emb_dim, self.num_classes or 0)
        h = C.c_void_p()
        _lib.check(lib.cdm_unet_create(C.byref(cfg), device_index, C.byref(h)))
        return h

    def _after_upload(self, lib, handle):
        # the sinusoidal frequency table, evaluated exactly like SinusoidalPosEmb.forward does
        half = self.base_dim // 2
        freq = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1))).float().contiguous()
        _lib.check(lib.cdm_unet_set_param(handle, b"@sin_freq", C.c_void_p(freq.data_ptr()), freq.numel()))

    # -- forward ---------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, t, y=None, precision=None):
        if self.num_classes is not None and y is None:
            raise ValueError("Class labels `y` must be provided for a conditional UNet.")
        _lib.require_cuda(x, t, y)
        self._inference_only()
        if x.dim() != 4 or x.shape[1] != self.in_channels or x.shape[2] != x.shape[3]:
            raise ValueError(f"expected x of shape [B, {self.in_channels}, S, S], got {tuple(x.shape)}")
        lib = _lib.lib()
        h = self._native_handle(x.device)
        prec = _lib.precision_code(precision or self.precision)
        B, S = x.shape[0], x.shape[2]
        x = x.detach().float().contiguous()
        t = t.detach().to(x.device, torch.float32).expand(B).contiguous()
        yy = y.detach().to(x.device, torch.int64).contiguous() if (y is not None and self.num_classes is not None) else None
        eps = torch.empty_like(x)
        with torch.cuda.device(x.device):
            nbytes = lib.cdm_unet_workspace_bytes(h, B, S, prec)
            ws = _native.workspace(x.device, nbytes)
            _lib.check(lib.cdm_unet_forward(h, _lib.ptr(x), _lib.ptr(t), _lib.ptr(yy), _lib.ptr(eps), B, S, prec,
                                            _lib.ptr(ws), ws.numel(), _lib.stream_of(x)))
        return eps

    @torch.no_grad()
    def forward_jvp(self, x, t, y, v_in, v_out=None, precision=None):
        """(eps, <v_out, J v_in>) per sample, J = d eps / d x, by forward-mode differentiation (the model's precision:
        fp32 CUDA-core convs, or fp16 tensor-core convs for primal and tangent).
        With v_out=None this is the Hutchinson term v^T J v of ``vector_field`` (shapes/compose_images_ito.py:46-63)."""
        if self.num_classes is not None and y is None:
            raise ValueError("Class labels `y` must be provided for a conditional UNet.")
        _lib.require_cuda(x, t, y, v_in, v_out)
        self._inference_only()
        lib = _lib.lib()
        h = self._native_handle(x.device)
        B, S = x.shape[0], x.shape[2]
        x = x.detach().float().contiguous()
        v_in = v_in.detach().float().contiguous()
        v_out = v_out.detach().float().contiguous() if v_out is not None else None
        t = t.detach().to(x.device, torch.float32).expand(B).contiguous()
        yy = y.detach().to(x.device, torch.int64).contiguous() if (y is not None and self.num_classes is not None) else None
        eps = torch.empty_like(x)
        vjv = torch.empty(B, device=x.device, dtype=torch.float32)
        prec = _lib.precision_code(precision or self.precision)
        with torch.cuda.device(x.device):
            nbytes = lib.cdm_unet_jvp_workspace_bytes(h, B, S, prec)
            ws = _native.workspace(x.device, nbytes)
            _lib.check(lib.cdm_unet_forward_jvp(h, _lib.ptr(x), _lib.ptr(t), _lib.ptr(yy), _lib.ptr(v_in), _lib.ptr(v_out),
                                                _lib.ptr(eps), _lib.ptr(vjv), B, S, prec, _lib.ptr(ws), ws.numel(),
                                                _lib.stream_of(x)))
        return eps, vjv

    def debug_read(self, name, B, S):
        """NCHW fp32 copy of an intermediate ("x0","d1","d2","b1","u1","u2") of the last forward."""
        lib = _lib.lib()
        d = self.base_dim
        shapes = {"x0": (d, S), "d1": (d, S), "d2": (2 * d, S // 2), "b1": (4 * d, S // 4), "u1": (2 * d, S // 2), "u2": (d, S)}
        c, s = shapes[name]
        dev = torch.device("cuda", torch.cuda.current_device())
        out = torch.empty(B, c, s, s, device=dev, dtype=torch.float32)
        _lib.check(lib.cdm_unet_debug_read(self._handle, name.encode(), _lib.ptr(out), B, S, _lib.stream_of(out)))
        return out


@torch.no_grad()
def forward_grouped(experts, xs, t, ys=None):
    """K native ``UNet`` experts evaluated with ONE grouped launch per convolution (``cdm_unet_forward_grouped``): what the
    reference does back to back (``eps_hat1 = model1(x, t); eps_hat2 = model2(x, t)``, mnist/compose_scores.py:33-34).
    xs: one tensor for all experts, or a list of K tensors (e.g. Grayscale(x) for a 1-channel expert next to x); ys: None
    or a list of K label tensors / None.  Returns the K predictions.  Bit-identical to calling the experts one by one."""
    lib = _lib.lib()
    K = len(experts)
    xs = [xs] * K if torch.is_tensor(xs) else list(xs)
    _lib.require_cuda(t, *xs)
    dev = xs[0].device
    B, S = xs[0].shape[0], xs[0].shape[2]
    for m, x in zip(experts, xs):
        m._inference_only()
        if x.dim() != 4 or x.shape[1] != m.in_channels or x.shape[2] != x.shape[3] or x.shape[0] != B or x.shape[2] != S:
            raise ValueError(f"expert with {m.in_channels} input channels got x of shape {tuple(x.shape)}")
    prec = _lib.precision_code(experts[0].precision)
    if any(_lib.precision_code(m.precision) != prec for m in experts):
        raise ValueError("forward_grouped: the experts must share one precision")
    xs = [x.detach().float().contiguous() for x in xs]
    t = t.detach().to(dev, torch.float32).expand(B).contiguous()
    ylist = []
    for k, m in enumerate(experts):
        y = ys[k] if ys is not None else None
        if m.num_classes is not None and y is None:
            raise ValueError("Class labels `y` must be provided for a conditional UNet.")
        ylist.append(y.detach().to(dev, torch.int64).contiguous() if (y is not None and m.num_classes is not None) else None)
    eps = [torch.empty_like(x) for x in xs]
    handles = (C.c_void_p * K)(*[m._native_handle(dev).value for m in experts])
    hp = C.cast(handles, C.POINTER(C.c_void_p))
    yarr = (C.c_void_p * K)(*[(y.data_ptr() if y is not None else None) for y in ylist])
    with torch.cuda.device(dev):
        ws = _native.workspace(dev, lib.cdm_unet_forward_grouped_workspace_bytes(hp, K, B, S, prec))
        _lib.check(lib.cdm_unet_forward_grouped(hp, K, _lib.ptr_array(xs), _lib.ptr(t), C.cast(yarr, C.POINTER(C.c_void_p)),
                                                _lib.ptr_array(eps), B, S, prec, _lib.ptr(ws), ws.numel(), _lib.stream_of(xs[0])))
    return eps
