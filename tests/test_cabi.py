"""CPU: the C-ABI shared library loads and exports every symbol include/dpf_sm100.h declares (no compute calls)."""
import ctypes
import re

import pytest

from conftest import ROOT
from dualpixelface_b200 import _lib

HEADER = ROOT / "include" / "dpf_sm100.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(dpf_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built():
    assert _lib.LIB_PATH.is_file(), "run `python __graft_entry__.py` (build()) first"


def test_every_declared_symbol_is_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 16
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes binding and header disagree"


def test_abi_version_and_cpu_refusal():
    lib = _lib.load()
    assert lib.dpf_abi_version() == 1
    import torch
    if not torch.cuda.is_available():
        assert lib.dpf_device_check() != 0                 # no sm_100a device: refuses, with a message
        assert len(lib.dpf_last_error()) > 0


def test_conv_struct_layout_matches_header():
    text = HEADER.read_text()
    body = text[text.index("typedef struct dpf_conv3d_args"):text.index("} dpf_conv3d_args;")]
    names = re.findall(r"\b(?:int|const void\*|void\*|const float\*|float\*|float)\s+([^;]+);", body)
    fields = [n.strip() for grp in names for n in grp.split(",")]
    fields = [re.sub(r"\s*/\*.*", "", f).strip() for f in fields]
    assert fields == [f[0] for f in _lib.ConvArgs._fields_], (fields, [f[0] for f in _lib.ConvArgs._fields_])


def test_ops_refuse_cpu_tensors():
    import torch
    from dualpixelface_b200 import ops
    x = torch.zeros(1, 4, 4, 32, dtype=torch.bfloat16)
    with pytest.raises(_lib.DpfError):
        ops.costvol_fwd(x, x, [0, 1])


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No CPU or PyTorch fallback: without the built .so the binding raises instead of degrading."""
    from dualpixelface_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "libdpf_sm100.so")
    with pytest.raises(_lib.DpfError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under dualpixelface_b200/, src/, main.py may import it (bench.py only in its
    cpu_baseline / --impl reference / gpu_eager_oracle baseline legs, __graft_entry__ only in smoke())."""
    import re
    from conftest import ROOT
    pat = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    offenders = [str(p.relative_to(ROOT)) for base in ("dualpixelface_b200", "src") for p in (ROOT / base).rglob("*.py")
                 if pat.search(p.read_text())]
    offenders += [f for f in ("main.py",) if pat.search((ROOT / f).read_text())]
    assert offenders == []
    # bench.py: only inside its two stated-BASELINE legs (the CPU port and the oracle's torch code run eagerly on the GPU), never in
    # the functions that produce `value` / `e2e` / `train`
    bench = (ROOT / "bench.py").read_text()
    funcs = re.split(r"^def ", bench, flags=re.M)
    importing = sorted(f.split("(")[0] for f in funcs if pat.search(f))
    assert importing == ["cpu_reference_pairs_per_s", "gpu_eager_oracle_block"]
