"""ctypes binding of libcdm_b200.so (the C ABI declared in include/cdm_b200.h).

There is no CPU fallback: if the library is missing it is built (nvcc must be
present) and if it cannot be loaded, or no sm_100 device is present when a
compute entry point is called, the error is raised -- never swallowed.
"""
import ctypes as C
import os

import torch  # noqa: F401  (initialises CUDA libraries before the .so is mapped)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CDM_LIB_PATH") or os.path.join(HERE, "libcdm_b200.so")   # override: kernel-variant experiments (tools/)

MAX_EXPERTS = 8
PREC_FP32, PREC_F16, PREC_F16X3 = 0, 1, 4
OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_NOT_READY, ERR_WORKSPACE, ERR_KEY = 0, -1, -2, -3, -4, -5, -6


class CdmError(RuntimeError):
    pass


class Rng(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("step", C.c_uint64)]


class ProfEntry(C.Structure):
    _fields_ = [("name", C.c_char * 24), ("launches", C.c_longlong), ("ms", C.c_double), ("flops", C.c_double),
                ("bytes", C.c_double)]


class UNetConfig(C.Structure):
    _fields_ = [("in_channels", C.c_int), ("base_dim", C.c_int), ("time_emb_dim", C.c_int), ("num_classes", C.c_int)]


_f, _i, _vp, _fp = C.c_float, C.c_int, C.c_void_p, C.c_void_p
_pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); mirrors include/cdm_b200.h one to one
SIGNATURES = {
    "cdm_abi_version": (_i, []),
    "cdm_abi_stamp": (C.c_uint, []),
    "cdm_last_error": (C.c_char_p, []),
    "cdm_launch_count": (C.c_longlong, []),
    "cdm_prof_enable": (_i, [_i]),
    "cdm_prof_summary": (_i, [C.POINTER(ProfEntry), _i]),
    "cdm_prof_dump": (_i, []),
    "cdm_device_check": (_i, [_i]),
    "cdm_step_sde": (_i, [_fp, _pp, C.POINTER(_i), C.POINTER(_f), _i, _fp, C.POINTER(Rng), _f, _f, _f, _f, _fp, _i, _i, _i, _vp]),
    "cdm_step_ddim": (_i, [_fp, _pp, C.POINTER(_i), C.POINTER(_f), _i, _f, _f, _f, _f, _f, _fp, _fp, _i, _i, _i, _vp]),
    "cdm_step_ddpm_logq": (_i, [_fp, _pp, _i, _fp, C.POINTER(Rng), _fp, _i, _f, _f, _f, _f, _f, _f, _f, _fp, _fp, _i, _i, _i, _vp]),
    "cdm_step_ode_kappa": (_i, [_fp, _fp, _i, _fp, _fp, _fp, _f, _i, _f, _f, _f, _f, _f, _f, _f, _fp, _fp, _i, _i, _i, _vp]),
    "cdm_step_ode_kappa_k": (_i, [_fp, _pp, C.POINTER(_i), _pp, C.POINTER(_f), _i, _f, _f, _f, _f, _f, _fp, _fp, _i, _i, _i, _vp]),
    "cdm_step_cfg": (_i, [_fp, _pp, C.POINTER(_f), _i, _f, _i, _i, _f, _f, _f, _f, _fp, C.POINTER(Rng), _fp, _i, _i, _i, _vp]),
    "cdm_step_superdiff_solve": (_i, [_fp, _pp, _i, _i, _f, _f, _f, _f, _f, _f, _f, _f, _f, _fp, _fp, C.POINTER(Rng), _fp, _fp, _fp,
                                      _i, _i, _i, _vp]),
    "cdm_latent_decode": (_i, [_fp, _fp, _fp, _fp, _i, _i, _i, _vp]),
    "cdm_step_layout": (_i, [_fp, _pp, _i, _vp, _i, _f, _f, _f, _f, _f, _fp, C.POINTER(Rng), _fp, _i, _i, _i, _vp]),
    "cdm_grayscale": (_i, [_fp, _fp, _i, _i, _vp]),
    "cdm_fill_normal": (_i, [_fp, C.c_int64, C.POINTER(Rng), _vp]),
    "cdm_unet_create": (_i, [C.POINTER(UNetConfig), _i, _pp]),
    "cdm_unet_destroy": (None, [_vp]),
    "cdm_unet_set_param": (_i, [_vp, C.c_char_p, _fp, C.c_int64]),
    "cdm_unet_finalize": (_i, [_vp]),
    "cdm_unet_num_params": (_i, [_vp]),
    "cdm_unet_param_key": (C.c_char_p, [_vp, _i, C.POINTER(C.c_int64)]),
    "cdm_set_microbatch": (_i, [_i]),
    "cdm_set_option": (_i, [C.c_char_p, _i]),
    "cdm_unet_workspace_bytes": (C.c_size_t, [_vp, _i, _i, _i]),
    "cdm_unet_forward": (_i, [_vp, _fp, _fp, _fp, _fp, _i, _i, _i, _vp, C.c_size_t, _vp]),
    "cdm_unet_forward_grouped_workspace_bytes": (C.c_size_t, [_pp, _i, _i, _i, _i]),
    "cdm_unet_forward_grouped": (_i, [_pp, _i, _pp, _fp, _pp, _pp, _i, _i, _i, _vp, C.c_size_t, _vp]),
    "cdm_unet_sample_workspace_bytes": (C.c_size_t, [_pp, _i, _i, _i, _i]),
    "cdm_unet_sample_sde": (_i, [_pp, C.POINTER(_f), _i, _fp, _pp, _i, _fp, C.POINTER(Rng), C.POINTER(_f), _i, _f, _i, _i, _i, _vp,
                                C.c_size_t, _vp]),
    "cdm_unet_sample_ddim_workspace_bytes": (C.c_size_t, [_pp, _i, _i, _i, _i, _i]),
    "cdm_unet_sample_ddim": (_i, [_pp, C.POINTER(_f), _i, _f, _fp, _pp, _i, C.POINTER(_f), _i, _i, _i, _i, _i, _vp, C.c_size_t, _vp]),
    "cdm_unet_sample_ito_workspace_bytes": (C.c_size_t, [_pp, _i, _i, _i, _i]),
    "cdm_unet_sample_ito": (_i, [_pp, _i, _fp, _pp, _i, _pp, C.POINTER(Rng), C.POINTER(_f), _i, _f, _i, _i, _i, _vp, C.c_size_t, _vp]),
    "cdm_score_sample_superdiff_workspace_bytes": (C.c_size_t, [_pp, _i, _i, _i, _i]),
    "cdm_score_sample_superdiff": (_i, [_pp, _i, _fp, _fp, _i, _f, _f, _fp, C.POINTER(Rng), C.POINTER(_f), _i, _i, _f, _i, _i, _i, _vp,
                                       C.c_size_t, _vp]),
    "cdm_guided_sample_cfg_workspace_bytes": (C.c_size_t, [_vp, _i, _i, _i]),
    "cdm_guided_sample_cfg": (_i, [_vp, _fp, _i, _i, _f, _f, C.POINTER(_f), _i, _i, _i, _i, _vp, C.c_size_t, _vp]),
    "cdm_unet_jvp_workspace_bytes": (C.c_size_t, [_vp, _i, _i, _i]),
    "cdm_unet_forward_jvp": (_i, [_vp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _vp, C.c_size_t, _vp]),
    "cdm_unet_debug_read": (_i, [_vp, C.c_char_p, _fp, _i, _i, _vp]),
    "cdm_score_create": (_i, [_i, _i, _i, _pp]),
    "cdm_score_destroy": (None, [_vp]),
    "cdm_score_set_param": (_i, [_vp, C.c_char_p, _fp, C.c_int64]),
    "cdm_score_finalize": (_i, [_vp]),
    "cdm_score_workspace_bytes": (C.c_size_t, [_vp, _i, _i]),
    "cdm_score_forward": (_i, [_vp, _fp, _fp, _fp, _i, _i, _vp, C.c_size_t, _vp]),
    "cdm_score_workspace_bytes_prec": (C.c_size_t, [_vp, _i, _i, _i]),
    "cdm_score_forward_prec": (_i, [_vp, _fp, _fp, _fp, _i, _i, _i, _vp, C.c_size_t, _vp]),
    "cdm_vae_decoder_create": (_i, [_i, _i, _pp]),
    "cdm_vae_decoder_destroy": (None, [_vp]),
    "cdm_vae_decoder_set_param": (_i, [_vp, C.c_char_p, _fp, C.c_int64]),
    "cdm_vae_decoder_finalize": (_i, [_vp]),
    "cdm_vae_decoder_workspace_bytes": (C.c_size_t, [_vp, _i]),
    "cdm_vae_decode": (_i, [_vp, _fp, _fp, _i, _vp, C.c_size_t, _vp]),
    "cdm_quantize_u8": (_i, [_fp, _vp, C.c_int64, _vp]),
    "cdm_simple_unet_create": (_i, [_i, _i, _pp]),
    "cdm_simple_unet_destroy": (None, [_vp]),
    "cdm_simple_unet_set_param": (_i, [_vp, C.c_char_p, _fp, C.c_int64]),
    "cdm_simple_unet_finalize": (_i, [_vp]),
    "cdm_simple_unet_workspace_bytes": (C.c_size_t, [_vp, _i, _i]),
    "cdm_simple_unet_forward": (_i, [_vp, _fp, _fp, _fp, _fp, _i, _i, _vp, C.c_size_t, _vp]),
    "cdm_simple_unet_workspace_bytes_prec": (C.c_size_t, [_vp, _i, _i, _i]),
    "cdm_simple_unet_forward_prec": (_i, [_vp, _fp, _fp, _fp, _fp, _i, _i, _i, _vp, C.c_size_t, _vp]),
    "cdm_guided_create": (_i, [_i, _i, _i, _i, _pp]),
    "cdm_guided_destroy": (None, [_vp]),
    "cdm_guided_set_param": (_i, [_vp, C.c_char_p, _fp, C.c_int64]),
    "cdm_guided_finalize": (_i, [_vp]),
    "cdm_guided_workspace_bytes": (C.c_size_t, [_vp, _i, _i, _i]),
    "cdm_guided_forward": (_i, [_vp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _vp, C.c_size_t, _vp]),
    "cdm_mlp_create": (_i, [_i, _i, _i, _pp]),
    "cdm_mlp_destroy": (None, [_vp]),
    "cdm_mlp_set_param": (_i, [_vp, C.c_char_p, _fp, C.c_int64]),
    "cdm_mlp_finalize": (_i, [_vp]),
    "cdm_mlp_forward": (_i, [_vp, _fp, _fp, _fp, _i, _vp]),
    "cdm_mlp_forward_jvp": (_i, [_vp, _fp, _fp, _fp, _fp, _fp, _i, _vp]),
    "cdm_mlp_sample_sde": (_i, [_pp, C.POINTER(_f), _i, _fp, _fp, C.POINTER(Rng), _fp, _i, _f, _i, _vp]),
    "cdm_mlp_sample_sde_tc": (_i, [_pp, C.POINTER(_f), _i, _fp, _fp, C.POINTER(Rng), _fp, _i, _f, _i, _vp]),
    "cdm_debug_init_conv": (_i, [_fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _vp]),
    "cdm_debug_maxpool": (_i, [_fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _vp]),
    "cdm_debug_upcat": (_i, [_fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "cdm_debug_conv_t16": (_i, [_fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "cdm_debug_conv": (_i, [_fp, _fp, _fp, _i, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
}

_lib = None


def lib():
    """Load (building first if needed) and return the ctypes library."""
    global _lib
    if _lib is None:
        from . import build as B
        if not os.environ.get("CDM_LIB_PATH"):
            # cheap when up to date (a source fingerprint is compared with the stamp); a stale .so is rebuilt, never loaded.
            # Concurrent ranks serialise on a file lock inside build().
            B.build()
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        if handle.cdm_abi_version() != 1:
            raise CdmError("libcdm_b200.so ABI version mismatch")
        if not os.environ.get("CDM_LIB_PATH") and handle.cdm_abi_stamp() != B.abi_stamp():
            raise CdmError("libcdm_b200.so was built from another include/cdm_b200.h (ABI stamp mismatch); rebuild with "
                           "python -m composable_diffusion_models_b200.build --force")
        _lib = handle
    return _lib


def check(status):
    if status == OK:
        return
    msg = lib().cdm_last_error().decode()
    if status == ERR_INVALID:
        raise ValueError(msg)
    if status == ERR_KEY:
        raise KeyError(msg)
    if status == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise CdmError(f"libcdm_b200 status {status}: {msg}")


def ptr(t):
    """Device pointer of a contiguous torch tensor (or None)."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise ValueError("libcdm_b200 needs contiguous tensors")
    return C.c_void_p(t.data_ptr())


def ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        if not t.is_contiguous():
            raise ValueError("libcdm_b200 needs contiguous tensors")
        arr[i] = t.data_ptr()
    return C.cast(arr, _pp)


def farray(vals):
    return (C.c_float * len(vals))(*[float(v) for v in vals])


def iarray(vals):
    return (C.c_int * len(vals))(*[int(v) for v in vals])


def stream_of(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise CdmError("composable_diffusion_models_b200 has no CPU path: tensors must live on a B200 (cuda) device")


def launch_count():
    return int(lib().cdm_launch_count())


def prof_enable(on=True):
    check(lib().cdm_prof_enable(1 if on else 0))


def prof_summary():
    """{kernel class: dict(launches, ms, flops, bytes)} since prof_enable(True); synchronises."""
    arr = (ProfEntry * 16)()
    n = lib().cdm_prof_summary(arr, 16)
    if n < 0:
        check(n)
    return {arr[i].name.decode(): dict(launches=int(arr[i].launches), ms=arr[i].ms, flops=arr[i].flops, bytes=arr[i].bytes)
            for i in range(n) if arr[i].launches}


def precision_code(p):
    if p in (PREC_FP32, "fp32", "float32", torch.float32):
        return PREC_FP32
    if p in (PREC_F16, "fp16", "f16", "float16", "half", torch.float16):
        return PREC_F16
    if p in (PREC_F16X3, "f16x3", "fp16x3", "x3"):
        return PREC_F16X3
    if p in ("bf16", "bfloat16", torch.bfloat16):
        raise ValueError("the tensor-core path computes with fp16 operands (fp32 accumulation): same tcgen05 rate as "
                         "bf16 at 1/8 of the rounding error; ask for precision='fp16'")
    raise ValueError(f"unknown precision {p!r}")
