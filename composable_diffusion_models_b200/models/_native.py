"""Shared plumbing for native experts: parameter upload and workspace caching."""
import ctypes as C

import torch

from .. import _lib

_workspaces = {}


def workspace(device, nbytes):
    """A per-device scratch buffer owned by torch's allocator, grown on demand."""
    key = (device.type, device.index)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        _workspaces[key] = buf = torch.empty(int(nbytes * 1.05) + 1024, dtype=torch.uint8, device=device)
    return buf


def param_signature(module):
    return tuple((k, v.data_ptr(), v._version, tuple(v.shape)) for k, v in module.state_dict().items())


def upload_state_dict(set_param, handle, state_dict):
    for key, value in state_dict.items():
        host = value.detach().to("cpu", torch.float32).contiguous()
        _lib.check(set_param(handle, key.encode(), C.c_void_p(host.data_ptr()), host.numel()))
