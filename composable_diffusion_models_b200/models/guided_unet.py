"""GuidedUNet (cross-attention UNet with digit + colour guidance), B200-native.

Same constructor, ``forward(x, t, digit_labels, color_labels)``, ``null_digit_idx`` / ``null_color_idx`` and
``state_dict()`` keys (including ``nn.MultiheadAttention``'s packed / unpacked projection parameters) as the
reference's ``src/compositional_diffusion_with_cross_attention.py:144-208``.  Each block attends to a context of
length ONE, so its softmax is identically 1 and the attention reduces to ``out_proj(v_proj(context))`` broadcast
over pixels; the native path folds that (``cdm_guided_finalize``) and never runs an attention kernel.

``precision``: "fp16" (every 3x3 conv and both ConvTranspose2d on tcgen05, default) or "fp32" (CUDA-core path that meets
the <= 1e-5 parity bound); the ``CDM_PRECISION`` environment variable overrides the default.
"""
import ctypes as C
import os

import torch
import torch.nn as nn

from .. import _lib
from . import _native


def _unet_block(in_channels, out_channels, time_emb_dim, context_dim):
    # registration order of the reference's UNetBlock
    b = nn.Module()
    b.time_mlp = nn.Linear(time_emb_dim, out_channels)
    b.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1)
    b.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1)
    b.norm1 = nn.GroupNorm(8, out_channels)
    b.norm2 = nn.GroupNorm(8, out_channels)
    b.attn = nn.Module()
    b.attn.attention = nn.MultiheadAttention(embed_dim=out_channels, kdim=context_dim, vdim=context_dim, num_heads=4,
                                             batch_first=True)
    b.attn_norm = nn.LayerNorm(out_channels)
    return b


class GuidedUNet(_native.NativeModule):
    _abi = "cdm_guided"

    def __init__(self, num_digits=10, num_colors=3, embed_dim=128, precision=None):
        super().__init__()
        self.embed_dim, self.num_digits, self.num_colors = embed_dim, num_digits, num_colors
        self.precision = precision or os.environ.get("CDM_PRECISION", "fp16")
        self.digit_embedding = nn.Embedding(num_digits + 1, embed_dim)
        self.color_embedding = nn.Embedding(num_colors + 1, embed_dim)
        self.null_digit_idx = num_digits
        self.null_color_idx = num_colors
        self.time_mlp = nn.ModuleDict({"1": nn.Linear(embed_dim, embed_dim)})
        self.init_conv = nn.Conv2d(3, 64, kernel_size=3, padding=1)
        cd = embed_dim * 2
        self.down1 = _unet_block(64, 128, embed_dim, cd)
        self.down2 = _unet_block(128, 256, embed_dim, cd)
        self.bot1 = _unet_block(256, 512, embed_dim, cd)
        self.bot2 = _unet_block(512, 256, embed_dim, cd)
        self.up1 = nn.ConvTranspose2d(256, 128, 2, 2)
        self.up2 = _unet_block(256 + 128, 128, embed_dim, cd)
        self.up3 = nn.ConvTranspose2d(128, 64, 2, 2)
        self.up4 = _unet_block(128 + 64, 64, embed_dim, cd)
        self.out_conv = nn.Conv2d(128, 3, kernel_size=1)

    def _create_native(self, lib, device_index):
        h = C.c_void_p()
        _lib.check(lib.cdm_guided_create(self.num_digits, self.num_colors, self.embed_dim, device_index, C.byref(h)))
        return h

    @torch.no_grad()
    def forward(self, x, t, digit_labels, color_labels):
        _lib.require_cuda(x, t, digit_labels, color_labels)
        self._inference_only()
        lib = _lib.lib()
        h = self._native_handle(x.device)
        B, S = x.shape[0], x.shape[2]
        x = x.detach().float().contiguous()
        t = t.detach().to(x.device, torch.float32).expand(B).contiguous()
        d = digit_labels.detach().to(x.device, torch.int64).expand(B).contiguous()
        c = color_labels.detach().to(x.device, torch.int64).expand(B).contiguous()
        eps = torch.empty_like(x)
        prec = _lib.precision_code(self.precision)
        with torch.cuda.device(x.device):
            ws = _native.workspace(x.device, lib.cdm_guided_workspace_bytes(h, B, S, prec))
            _lib.check(lib.cdm_guided_forward(h, _lib.ptr(x), _lib.ptr(t), _lib.ptr(d), _lib.ptr(c), _lib.ptr(eps), B, S, prec,
                                              _lib.ptr(ws), ws.numel(), _lib.stream_of(x)))
        return eps
