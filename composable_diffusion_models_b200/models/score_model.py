"""ColoredMNISTScoreModel / ScoreModel (BatchNorm UNet of the SuperDiff scripts), B200-native.

Same constructor, ``forward(x, t)`` and ``state_dict()`` keys (BatchNorm buffers included) as the reference's
``src/models/compose_grayscale_object_and_color.py:80-112`` (identical class in
``src/models/composing_colored_digit_to_simulate_overlaying.py``).  Sampling uses eval-mode BatchNorm (running
statistics), as ``SuperDiffSampler.sample`` puts the experts in ``.eval()`` (src/diffusion/samplers.py:13-14).
The modules below only hold parameters; the forward pass is ``cdm_score_forward`` (fp32 path).
"""
import ctypes as C
import os

import torch
import torch.nn as nn

from .. import _lib
from . import _native


def _block(in_ch, out_ch, time_emb_dim, up=False, transform=True):
    # registration order of the reference's Block / ConvBlock (time_mlp, conv1, transform, conv2, bnorm1, bnorm2)
    b = nn.Module()
    b.time_mlp = nn.Linear(time_emb_dim, out_ch)
    b.conv1 = nn.Conv2d(2 * in_ch if up else in_ch, out_ch, 3, padding=1)
    if transform:
        b.transform = nn.Conv2d(out_ch, out_ch, 4, 2, 1)
    b.conv2 = nn.Conv2d(out_ch, out_ch, 3, padding=1)
    b.bnorm1 = nn.BatchNorm2d(out_ch)
    b.bnorm2 = nn.BatchNorm2d(out_ch)
    return b


class ColoredMNISTScoreModel(_native.NativeModule):
    _abi = "cdm_score"

    def __init__(self, in_channels: int = 3, time_emb_dim: int = 32, precision=None):
        super().__init__()
        self.in_channels, self.time_emb_dim = in_channels, time_emb_dim
        self.precision = precision or os.environ.get("CDM_PRECISION", "fp16")
        self.time_mlp = nn.ModuleDict({"1": nn.Linear(time_emb_dim, time_emb_dim * 4),
                                       "3": nn.Linear(time_emb_dim * 4, time_emb_dim)})
        self.initial_conv = nn.Conv2d(in_channels, 32, 3, padding=1)
        self.down1 = _block(32, 64, time_emb_dim)
        self.down2 = _block(64, 128, time_emb_dim)
        self.bot1 = _block(128, 256, time_emb_dim)
        self.up_transpose_1 = nn.ConvTranspose2d(256, 128, 4, 2, 1)
        self.up_block_1 = _block(256, 128, time_emb_dim, transform=False)
        self.up_transpose_2 = nn.ConvTranspose2d(128, 64, 4, 2, 1)
        self.up_block_2 = _block(128, 64, time_emb_dim, transform=False)
        self.up_transpose_3 = nn.ConvTranspose2d(64, 32, 4, 2, 1)
        self.up_block_3 = _block(64, 32, time_emb_dim, transform=False)
        self.output = nn.Conv2d(32, in_channels, 1)

    def _create_native(self, lib, device_index):
        h = C.c_void_p()
        _lib.check(lib.cdm_score_create(self.in_channels, self.time_emb_dim, device_index, C.byref(h)))
        return h

    @torch.no_grad()
    def forward(self, x, t):
        _lib.require_cuda(x, t)
        if self.training:
            raise NotImplementedError("libcdm_b200 implements the sampling (eval-mode BatchNorm) path only; call .eval()")
        lib = _lib.lib()
        h = self._native_handle(x.device)
        B, S = x.shape[0], x.shape[2]
        x = x.detach().float().contiguous()
        t = t.detach().to(x.device, torch.float32).expand(B).contiguous()
        eps = torch.empty_like(x)
        prec = _lib.precision_code(self.precision)
        with torch.cuda.device(x.device):
            ws = _native.workspace(x.device, lib.cdm_score_workspace_bytes_prec(h, B, S, prec))
            _lib.check(lib.cdm_score_forward_prec(h, _lib.ptr(x), _lib.ptr(t), _lib.ptr(eps), B, S, prec, _lib.ptr(ws), ws.numel(),
                                                  _lib.stream_of(x)))
        return eps


ScoreModel = ColoredMNISTScoreModel
