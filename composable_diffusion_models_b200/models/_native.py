"""Shared plumbing for native experts: the library handle's life cycle, parameter upload and workspace caching."""
import ctypes as C
import itertools
import warnings

import torch
import torch.nn as nn

from .. import _lib

_workspaces = {}


def workspace(device, nbytes):
    """A scratch buffer owned by torch's allocator, grown on demand, one per (device, CUDA stream): experts driven
    from different streams never share activations."""
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        _workspaces[key] = buf = torch.empty(int(nbytes * 1.05) + 1024, dtype=torch.uint8, device=device)
    return buf


def param_signature(module):
    """Changes whenever a parameter / buffer is replaced, moved or written in place (cheap: no state_dict() build)."""
    return tuple((v.data_ptr(), v._version) for v in itertools.chain(module.parameters(), module.buffers()))


def upload_state_dict(set_param, handle, state_dict):
    for key, value in state_dict.items():
        host = value.detach().to("cpu", torch.float32).contiguous()
        _lib.check(set_param(handle, key.encode(), C.c_void_p(host.data_ptr()), host.numel()))


class NativeModule(nn.Module):
    """A parameter holder whose forward runs in libcdm_b200.  Sub-classes set ``_abi`` (the C prefix, e.g. "cdm_unet")
    and implement ``_create_native(lib, device_index) -> c_void_p``; ``_after_upload(lib, handle)`` is optional.

    The native handle packs the weights on ONE device.  It is rebuilt when the module moves to another device, re-packed
    when a parameter changes, never shared by copies (``copy.deepcopy`` / pickling drop it) and freed with the module."""

    _abi = None
    _train_note = "its train-mode behaviour (dropout / batch statistics) is not implemented"

    def __init__(self):
        super().__init__()
        self._handle = None
        self._handle_dev = None
        self._sig = None

    def _create_native(self, lib, device_index):
        raise NotImplementedError

    def _after_upload(self, lib, handle):
        pass

    def _fn(self, lib, what):
        return getattr(lib, f"{self._abi}_{what}")

    def _release_native(self):
        h, self._handle, self._handle_dev, self._sig = self._handle, None, None, None
        if h is not None:
            self._fn(_lib.lib(), "destroy")(h)

    def _native_handle(self, device):
        lib = _lib.lib()
        dev = device.index if device.index is not None else torch.cuda.current_device()
        if self._handle is not None and self._handle_dev != dev:
            self._release_native()          # weights live on the old device: a fresh handle, not a re-upload into it
        sig = param_signature(self)
        if self._handle is not None and sig == self._sig:
            return self._handle
        if self._handle is None:
            self._handle = self._create_native(lib, dev)
            self._handle_dev = dev
        upload_state_dict(self._fn(lib, "set_param"), self._handle, self.state_dict())
        self._after_upload(lib, self._handle)
        with torch.cuda.device(dev):
            _lib.check(self._fn(lib, "finalize")(self._handle))
        self._sig = sig
        return self._handle

    def _inference_only(self):
        """The native path is the sampling path: no dropout, eval-mode normalisation, no autograd graph."""
        if self.training and not getattr(self, "_warned_training", False):
            self.__dict__["_warned_training"] = True
            warnings.warn(f"{type(self).__name__} runs the inference path of libcdm_b200 although the module is in training "
                          f"mode ({self._train_note}); call .eval() as the reference samplers do", stacklevel=3)

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_handle"] = state["_handle_dev"] = state["_sig"] = None
        return state

    def __del__(self):
        try:
            self._release_native()
        except Exception:
            pass
