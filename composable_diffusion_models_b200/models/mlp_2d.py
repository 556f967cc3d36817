"""The 2-D latent MLP expert, B200-native.

Same constructor, ``forward(t, x)`` argument order and ``state_dict()`` keys (``main.{0,2,4,6}``) as the
reference's ``mnist/models/mlp_2d.py:5-20`` (== ``shapes/models/mlp_2d.py``).
"""
import ctypes as C

import torch
import torch.nn as nn

from .. import _lib
from . import _native


class MLP(_native.NativeModule):
    _abi = "cdm_mlp"

    def __init__(self, num_hid=256, num_out=2):
        super().__init__()
        self.num_hid, self.num_out = num_hid, num_out
        self.main = nn.ModuleDict({
            "0": nn.Linear(1 + num_out, num_hid),
            "2": nn.Linear(num_hid, num_hid),
            "4": nn.Linear(num_hid, num_hid),
            "6": nn.Linear(num_hid, num_out),
        })

    def _create_native(self, lib, device_index):
        h = C.c_void_p()
        _lib.check(lib.cdm_mlp_create(self.num_hid, self.num_out, device_index, C.byref(h)))
        return h

    @torch.no_grad()
    def forward(self, t, x):
        _lib.require_cuda(t, x)
        lib = _lib.lib()
        h = self._native_handle(x.device)
        B = x.shape[0]
        x = x.detach().float().contiguous()
        t = t.detach().to(x.device, torch.float32).reshape(-1).expand(B).contiguous()
        eps = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(lib.cdm_mlp_forward(h, _lib.ptr(t), _lib.ptr(x), _lib.ptr(eps), B, _lib.stream_of(x)))
        return eps

    @torch.no_grad()
    def forward_jvp(self, t, x, v):
        """(eps, v^T J v) per sample by forward-mode differentiation (``vector_field`` of the latent Ito scripts)."""
        _lib.require_cuda(t, x, v)
        lib = _lib.lib()
        h = self._native_handle(x.device)
        B = x.shape[0]
        x = x.detach().float().contiguous()
        v = v.detach().float().contiguous()
        t = t.detach().to(x.device, torch.float32).reshape(-1).expand(B).contiguous()
        eps = torch.empty_like(x)
        vjv = torch.empty(B, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(lib.cdm_mlp_forward_jvp(h, _lib.ptr(t), _lib.ptr(x), _lib.ptr(v), _lib.ptr(eps), _lib.ptr(vjv), B,
                                               _lib.stream_of(x)))
        return eps, vjv
