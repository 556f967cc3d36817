"""SimpleUnet, the 62 M-parameter GroupNorm UNet of the classifier-free-guidance SuperDiff scripts, B200-native.

Same constructor, ``forward(x, timestep, y)`` and ``state_dict()`` keys as the reference's ``SimpleUnet``
(``src/composing_conditional_diffusion_on_shape_and_color_6.py:184-221``; identical in ``_6_1`` and ``_7``): label ids in
``[0, num_classes]`` with ``num_classes`` the null / unconditional token.  The modules hold parameters; the forward pass is
``cdm_simple_unet_forward`` (fp32 path).
"""
import ctypes as C
import os

import torch
import torch.nn as nn

from .. import _lib
from . import _native

_DOWN = (64, 128, 256, 512, 1024)
_UP = (1024, 512, 256, 128, 64)


def _block(in_ch, out_ch, time_emb_dim, up=False):
    # registration order of the reference's Block: time_mlp, conv1, transform, conv2, gn1, gn2
    b = nn.Module()
    b.time_mlp = nn.Linear(time_emb_dim, out_ch)
    if up:
        b.conv1 = nn.Conv2d(2 * in_ch, out_ch, 3, padding=1)
        b.transform = nn.ConvTranspose2d(out_ch, out_ch, 4, 2, 1)
    else:
        b.conv1 = nn.Conv2d(in_ch, out_ch, 3, padding=1)
        b.transform = nn.Conv2d(out_ch, out_ch, 4, 2, 1)
    b.conv2 = nn.Conv2d(out_ch, out_ch, 3, padding=1)
    b.gn1 = nn.GroupNorm(8, out_ch)
    b.gn2 = nn.GroupNorm(8, out_ch)
    return b


class SimpleUnet(_native.NativeModule):
    _abi = "cdm_simple_unet"

    def __init__(self, num_classes, precision=None):
        super().__init__()
        self.num_classes = num_classes
        self.precision = precision or os.environ.get("CDM_PRECISION", "fp16")
        td = 32
        self.time_mlp = nn.ModuleDict({"1": nn.Linear(td, td)})
        self.label_emb = nn.Embedding(num_classes + 1, td)
        self.conv0 = nn.Conv2d(3, _DOWN[0], 3, padding=1)
        self.downs = nn.ModuleList([_block(_DOWN[i], _DOWN[i + 1], td) for i in range(4)])
        self.ups = nn.ModuleList([_block(_UP[i], _UP[i + 1], td, up=True) for i in range(4)])
        self.output = nn.Conv2d(_UP[-1], 3, 1)

    def _create_native(self, lib, device_index):
        h = C.c_void_p()
        _lib.check(lib.cdm_simple_unet_create(self.num_classes, device_index, C.byref(h)))
        return h

    @torch.no_grad()
    def forward(self, x, timestep, y):
        _lib.require_cuda(x, timestep, y)
        self._inference_only()
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != x.shape[3]:
            raise ValueError(f"expected x of shape [B, 3, S, S], got {tuple(x.shape)}")
        lib = _lib.lib()
        h = self._native_handle(x.device)
        B, S = x.shape[0], x.shape[2]
        x = x.detach().float().contiguous()
        t = timestep.detach().to(x.device, torch.float32).expand(B).contiguous()
        yy = y.detach().to(x.device, torch.int64).expand(B).contiguous()
        eps = torch.empty_like(x)
        prec = _lib.precision_code(self.precision)
        with torch.cuda.device(x.device):
            ws = _native.workspace(x.device, lib.cdm_simple_unet_workspace_bytes_prec(h, B, S, prec))
            _lib.check(lib.cdm_simple_unet_forward_prec(h, _lib.ptr(x), _lib.ptr(t), _lib.ptr(yy), _lib.ptr(eps), B, S, prec, _lib.ptr(ws),
                                                        ws.numel(), _lib.stream_of(x)))
        return eps
