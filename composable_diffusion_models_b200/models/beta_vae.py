"""BetaVAE decoder (the image-space epilogue of the latent-diffusion samplers), B200-native.

Same constructor argument, ``decode(z)`` and ``state_dict()`` keys as the reference's ``BetaVAE``
(``src/4.3 best_of_both_worlds_3.py:95-126``); ``sample_composed_latent`` (``:262-293``) ends with ``vae_decoder(z)`` where
``vae_decoder = vae.decode``.  Only the decoder half is on the sampling path: the encoder / fc_mu / fc_log_var modules are
registered (so a trained checkpoint loads with ``strict=True``) but ``forward`` / ``encode`` are not implemented.
The modules hold parameters; the computation is ``cdm_vae_decode`` (fp32 path).
"""
import ctypes as C

import torch
import torch.nn as nn

from .. import _lib
from . import _native


class BetaVAE(_native.NativeModule):
    _abi = "cdm_vae_decoder"

    def __init__(self, latent_dims: int):
        super().__init__()
        self.latent_dims = latent_dims
        # registration order and Sequential indices of the reference (ReLU / Flatten / Unflatten / Sigmoid hold no parameters)
        self.encoder = nn.ModuleDict({"0": nn.Conv2d(3, 32, 4, 2, 1), "2": nn.Conv2d(32, 64, 4, 2, 1),
                                      "4": nn.Conv2d(64, 128, 4, 2, 1), "7": nn.Linear(128 * 4 * 4, 256)})
        self.fc_mu = nn.Linear(256, latent_dims)
        self.fc_log_var = nn.Linear(256, latent_dims)
        self.decoder_input = nn.Linear(latent_dims, 256)
        self.decoder = nn.ModuleDict({"0": nn.Linear(256, 128 * 4 * 4), "3": nn.ConvTranspose2d(128, 64, 4, 2, 1),
                                      "5": nn.ConvTranspose2d(64, 32, 4, 2, 1), "7": nn.ConvTranspose2d(32, 3, 4, 2, 1)})

    def _create_native(self, lib, device_index):
        h = C.c_void_p()
        _lib.check(lib.cdm_vae_decoder_create(self.latent_dims, device_index, C.byref(h)))
        return h

    @torch.no_grad()
    def decode(self, z):
        _lib.require_cuda(z)
        if z.dim() != 2 or z.shape[1] != self.latent_dims:
            raise ValueError(f"decode expects z of shape [B, {self.latent_dims}], got {tuple(z.shape)}")
        lib = _lib.lib()
        h = self._native_handle(z.device)
        B = z.shape[0]
        z = z.detach().float().contiguous()
        out = torch.empty(B, 3, 32, 32, device=z.device, dtype=torch.float32)
        with torch.cuda.device(z.device):
            ws = _native.workspace(z.device, lib.cdm_vae_decoder_workspace_bytes(h, B))
            _lib.check(lib.cdm_vae_decode(h, _lib.ptr(z), _lib.ptr(out), B, _lib.ptr(ws), ws.numel(), _lib.stream_of(z)))
        return out

    def forward(self, x):
        raise NotImplementedError("libcdm_b200 implements BetaVAE.decode (the sampling path) only")

    encode = forward


def quantize_u8(images):
    """``torchvision.utils.save_image``'s quantisation of a [0, 1] image tensor: uint8(clamp(x*255 + 0.5, 0, 255))."""
    _lib.require_cuda(images)
    x = images.detach().float().contiguous()
    out = torch.empty(x.shape, device=x.device, dtype=torch.uint8)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().cdm_quantize_u8(_lib.ptr(x), _lib.ptr(out), x.numel(), _lib.stream_of(x)))
    return out
