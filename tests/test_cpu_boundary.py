"""CPU: the C-ABI library builds/loads, exports every symbol include/cdm_b200.h declares, and the host-side
mirrors keep the reference's contracts (state_dict keys, error behaviour, schedule API).  No compute calls."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from composable_diffusion_models_b200 import _lib, schedule
from composable_diffusion_models_b200.models import MLP, UNet
from oracle import experts as E
from oracle import schedule as OS


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "cdm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cdm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"libcdm_b200.so does not export {n}"
    assert set(names) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert lib.cdm_abi_version() == 1


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, never compute on the host."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert _lib.lib().cdm_device_check(0) != 0
    assert b"CUDA" in _lib.lib().cdm_last_error() or b"device" in _lib.lib().cdm_last_error()
    with pytest.raises(_lib.CdmError):
        UNet()(torch.zeros(1, 1, 28, 28), torch.zeros(1))
    with pytest.raises(_lib.CdmError):
        MLP()(torch.zeros(4), torch.zeros(4, 2))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "composable_diffusion_models_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|from\s+\.+\s*import\s+oracle|oracle/", src, re.M), f"{f} uses the oracle"


@pytest.mark.parametrize("kw", [dict(in_channels=1), dict(in_channels=1, num_classes=3), dict(in_channels=3, num_classes=3)])
def test_unet_state_dict_contract(kw):
    m = UNet(**kw)
    spec = E.unet_small_spec(**kw)
    sd = m.state_dict()
    assert list(sd.keys()) == list(spec.keys())
    for k, shp in spec.items():
        assert tuple(sd[k].shape) == tuple(shp), k
    m.load_state_dict(E.synth_state_dict(spec, 3), strict=True)
    with pytest.raises(RuntimeError):
        bad = dict(E.synth_state_dict(spec, 3))
        bad.pop("out_conv.bias")
        m.load_state_dict(bad, strict=True)


def test_native_param_table_matches_state_dict():
    """The C side's expected key list (cdm_unet_param_key) equals the reference's state_dict key set."""
    lib = _lib.lib()
    for kw in (dict(in_channels=1, num_classes=None), dict(in_channels=3, num_classes=3)):
        cfg = _lib.UNetConfig(kw["in_channels"], 64, 256, kw["num_classes"] or 0)
        h = ctypes.c_void_p()
        _lib.check(lib.cdm_unet_create(ctypes.byref(cfg), 0, ctypes.byref(h)))
        n = lib.cdm_unet_num_params(h)
        got = {}
        for i in range(n):
            numel = ctypes.c_int64()
            key = lib.cdm_unet_param_key(h, i, ctypes.byref(numel)).decode()
            got[key] = numel.value
        spec = E.unet_small_spec(kw["in_channels"], num_classes=kw["num_classes"])
        want = {k: int(torch.Size(s).numel()) for k, s in spec.items()}
        assert got == want
        buf = torch.zeros(5)
        assert lib.cdm_unet_set_param(h, b"no.such.key", ctypes.c_void_p(buf.data_ptr()), 5) == _lib.ERR_KEY
        assert lib.cdm_unet_set_param(h, b"out_conv.bias", ctypes.c_void_p(buf.data_ptr()), 5) == _lib.ERR_KEY
        assert lib.cdm_unet_finalize(h) == _lib.ERR_KEY      # missing keys
        assert b"missing key" in lib.cdm_last_error()
        lib.cdm_unet_destroy(h)


def test_score_and_guided_state_dict_contracts():
    from composable_diffusion_models_b200.models import ColoredMNISTScoreModel, GuidedUNet
    m = ColoredMNISTScoreModel()
    spec = E.score_model_spec()
    assert list(m.state_dict().keys()) == list(spec.keys())
    m.load_state_dict(E.synth_state_dict(spec, 2), strict=True)
    g = GuidedUNet()
    spec = E.guided_unet_spec()
    assert set(g.state_dict().keys()) == set(spec.keys())
    assert all(tuple(g.state_dict()[k].shape) == tuple(v) for k, v in spec.items())
    g.load_state_dict(E.synth_state_dict(spec, 2), strict=True)


def test_mlp_state_dict_contract():
    m = MLP()
    spec = E.mlp_2d_spec()
    assert list(m.state_dict().keys()) == list(spec.keys())
    m.load_state_dict(E.synth_state_dict(spec, 1), strict=True)


def test_conditional_unet_requires_labels():
    with pytest.raises(ValueError):
        UNet(in_channels=1, num_classes=3)(torch.zeros(1, 1, 28, 28), torch.zeros(1))


def test_schedule_api_matches_oracle():
    t = torch.linspace(1e-3, 1.0, 33)
    for n in ("log_alpha", "alpha", "sigma", "dlog_alphadt", "beta", "g2"):
        assert torch.equal(getattr(schedule, n)(t), getattr(OS, n)(t)), n
        assert getattr(schedule, n)(0.3).dtype == torch.float32          # floats are accepted, fp32 comes back
    assert torch.equal(schedule.jax_faithful.sigma(t), OS.jax_sigma(t))
    assert torch.equal(schedule.jax_faithful.beta(t), OS.jax_beta(t))
    assert torch.equal(schedule.jax_faithful.g2(t), OS.jax_g2(t))
    sde, o = schedule.VPSDE(), OS.VPSDETables()
    for n in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_one_minus_alphas_cumprod", "posterior_variance"):
        assert torch.equal(getattr(sde, n), getattr(o, n)), n
    x0 = torch.randn(4, 1, 8, 8)
    eps = torch.randn_like(x0)
    tt = torch.rand(4)
    assert torch.equal(schedule.q_t(x0, tt, eps)[0], OS.q_t(x0, tt, eps)[0])


def test_checkpoint_formats(tmp_path):
    from composable_diffusion_models_b200.utils import CheckpointManager, load_checkpoint, save_checkpoint
    m = UNet()
    sd = E.synth_state_dict(E.unet_small_spec(1), 9)
    m.load_state_dict(sd)
    save_checkpoint(m, None, 3, str(tmp_path / "a" / "fmtA.pth"))
    m2 = UNet()
    assert load_checkpoint(m2, None, str(tmp_path / "a" / "fmtA.pth"), "cpu") == 3
    assert all(torch.equal(m2.state_dict()[k], sd[k]) for k in sd)
    mgr = CheckpointManager(tmp_path, "exp", "run")
    mgr.save(m, "expert")
    m3 = mgr.load(UNet(), "expert", "cpu")
    assert all(torch.equal(m3.state_dict()[k], sd[k]) for k in sd)
    with pytest.raises(FileNotFoundError):
        mgr.load(UNet(), "missing", "cpu")
    torch.save(sd, tmp_path / "fmtB.pth")
    load_checkpoint(UNet(), None, str(tmp_path / "fmtB.pth"), "cpu")
