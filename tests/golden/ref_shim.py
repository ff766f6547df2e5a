"""Import the UNMODIFIED reference from /root/reference through a thin shim layer.

Only used here, in the build container, by ``make_golden.py`` and by the optional
``tests/test_oracle_vs_reference.py`` (skipped when /root/reference is absent, e.g. on the GPU box).
The shims do not change any arithmetic of the reference; they only supply what its 2020-era
environment had: ``pytorch_lightning``, ``texttable``, a CUDA device, ``torch.rfft/irfft`` and the compiled
``DCN`` extension (restated by ``oracle.dpf_oracle.deform_conv3d`` -- the one piece that is not the
reference's own code, see oracle/__init__.py).
"""
from __future__ import annotations

import json
import os
import sys
import types
from contextlib import contextmanager
from pathlib import Path

import torch

REFERENCE_ROOT = Path(os.environ.get("DPF_REFERENCE_ROOT", "/root/reference"))
_REPO_ROOT = Path(__file__).resolve().parents[2]
_installed = False


def reference_available() -> bool:
    return (REFERENCE_ROOT / "src" / "model" / "stereodpnet" / "mainmodel.py").is_file()


class _Obj:
    def __init__(self, d):
        for k, v in d.items():
            setattr(self, k, _Obj(v) if isinstance(v, dict) else v)


def _install_shims():
    global _installed
    if _installed:
        return
    sys.dont_write_bytecode = True
    if str(_REPO_ROOT) not in sys.path:
        sys.path.insert(0, str(_REPO_ROOT))
    from oracle.dpf_oracle import deform_conv3d

    # 1. pytorch_lightning / texttable stand-ins
    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(torch.nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

    pl.LightningModule = LightningModule
    sys.modules["pytorch_lightning"] = pl
    tt = types.ModuleType("texttable")
    tt.Texttable = type("Texttable", (), {"HEADER": 0, "set_deco": lambda s, *a: None})
    sys.modules["texttable"] = tt

    # 2. CPU only: neutralise .cuda(); keep "zeros(requires_grad=True) -> non-leaf copy" of a GPU run
    torch.Tensor.cuda = lambda self, *a, **k: self
    _zeros = torch.zeros

    def zeros(*a, **k):
        rg = k.pop("requires_grad", False)
        z = _zeros(*a, **k)
        return z.requires_grad_(True).clone() if rg else z

    torch.zeros = zeros

    # 3. legacy FFT API (src/module/asm/asm.py:112,125)
    torch.rfft = lambda x, n, onesided=True: torch.view_as_real(torch.fft.fft2(x))

    def irfft(X, n, onesided=True):
        Xc = torch.view_as_complex(X.contiguous())
        H, W = Xc.shape[-2:]
        return torch.fft.irfft2(Xc[..., : W // 2 + 1], s=(H, W))

    torch.irfft = irfft

    # 4. compiled D3D op -> top-level package `functions` exposing DeformConvFunction.apply
    class DeformConvFunction:
        @staticmethod
        def apply(inp, offset, weight, bias, stride, padding, dilation, group, deformable_groups, im2col_step):
            s = stride[0] if isinstance(stride, (tuple, list)) else stride
            p = padding[0] if isinstance(padding, (tuple, list)) else padding
            d = dilation[0] if isinstance(dilation, (tuple, list)) else dilation
            assert group == 1 and deformable_groups == 1
            return deform_conv3d(inp, offset, weight, bias, s, p, d)

    fpkg = types.ModuleType("functions")
    fsub = types.ModuleType("functions.deform_conv_func")
    fsub.DeformConvFunction = DeformConvFunction
    fpkg.deform_conv_func = fsub
    fpkg.DeformConvFunction = DeformConvFunction
    sys.modules["functions"] = fpkg
    sys.modules["functions.deform_conv_func"] = fsub
    _installed = True


@contextmanager
def _in_reference_tree():
    old = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    sys.path.insert(0, str(REFERENCE_ROOT))
    try:
        yield
    finally:
        os.chdir(old)
        sys.path.remove(str(REFERENCE_ROOT))


def reference_option(model_name: str, **model_overrides):
    cfg = json.loads((REFERENCE_ROOT / "config_" / ("train_faceDP.json" if model_name == "stereodpnet"
                                                     else "train_faceDP_psmnet.json")).read_text())
    cfg["model"] = json.loads((REFERENCE_ROOT / "src" / "model" / model_name / "config.json").read_text())
    cfg["dataset"] = json.loads((REFERENCE_ROOT / "dataloader" / "FaceDP" / "config.json").read_text())
    cfg["model"]["metric_type"] = [m for m in cfg["model"]["metric_type"] if m != "affine_dp"]
    cfg["model"].update(model_overrides)
    cfg["load_model"] = None
    return _Obj(cfg)


def build_reference_model(model_name: str, **model_overrides):
    """Returns the reference's own LightningModule (STEREODPNET / PSMNET), constructed under the shims."""
    _install_shims()
    from runpy import run_path
    with _in_reference_tree():
        opt = reference_option(model_name, **model_overrides)
        ns = run_path(str(Path("src/model") / model_name / "mainmodel.py"))
        model = ns[model_name.upper()](opt)
    return model


def reference_modules():
    """The reference's stage modules (importable as-is once the shims are in)."""
    _install_shims()
    with _in_reference_tree():
        import importlib
        mods = {
            "asm": importlib.import_module("src.module.asm.asm"),
            "sdp": importlib.import_module("src.model.stereodpnet.modules"),
            "psm": importlib.import_module("src.model.psmnet.modules"),
        }
    return mods


def reset_shift_cache(model):
    """The reference caches its sampling grids forever (asm.py:29-30,56-57); clear between shapes."""
    sl = model.cost_volume.shifting_layer
    sl.basic_grid_forward = sl.basic_grid_backward = sl.phase_grid_forward = sl.phase_grid_backward = None
    # ANM likewise registers its pixel grid once, at the first resolution it sees (normal_module.py:91-99)
    anm = getattr(model, "normal_estimator", None)
    if anm is not None and getattr(anm, "grid_check", False):
        anm.grid_check = False
        anm._parameters.pop("grid", None)
