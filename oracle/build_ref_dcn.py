"""Builds the reference's own 3-D deformable-convolution CUDA extension into oracle/_ref/DCN.so (test infrastructure only).

The sources are compiled WHERE THEY LIE under /root/reference/src/module/dcn3d/src (nothing is copied into this repository);
the only addition is the force-included header oracle/ref_dcn_shim.h, which re-declares one dispatch macro for current PyTorch.
The resulting pybind module exposes DCN.deform_conv_forward / deform_conv_backward (src/vision.cpp:4-7, src/deform_conv.h:10-69)
and is used by tests/test_gpu_dcn_reference.py to pin both the oracle's restatement and the sm_100a kernels against the
reference's real kernels on a B200.  /root/reference does not exist on the GPU box: the prebuilt .so travels with the snapshot
(oracle/_ref/ is git-ignored, not gpurun-ignored).

    python oracle/build_ref_dcn.py        # no GPU needed (nvcc cross-compiles sm_100a)
"""
from __future__ import annotations

import glob
import os
import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = Path("/root/reference/src/module/dcn3d/src")
OUT = HERE / "_ref"


def build(verbose: bool = False) -> Path | None:
    if not SRC.is_dir():
        print(f"[oracle/_ref] {SRC} not present: keeping the prebuilt DCN.so (if any)")
        return None
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load
    bdir = OUT / "dcn_build"
    bdir.mkdir(parents=True, exist_ok=True)
    shim = str(HERE / "ref_dcn_shim.h")
    sources = [str(SRC / "vision.cpp")] + sorted(glob.glob(str(SRC / "cpu" / "*.cpp"))) + sorted(glob.glob(str(SRC / "cuda" / "*.cu")))
    load(name="DCN", sources=sources, extra_include_paths=[str(SRC)],
         extra_cflags=["-DWITH_CUDA", "-include", shim, "-w"],
         extra_cuda_cflags=["-DWITH_CUDA", "-include", shim, "-w", "-gencode", "arch=compute_100a,code=sm_100a"],
         build_directory=str(bdir), verbose=verbose, is_python_module=False)
    so = bdir / "DCN.so"
    shutil.copy2(so, OUT / "DCN.so")
    print(f"[oracle/_ref] built {OUT / 'DCN.so'} ({(OUT / 'DCN.so').stat().st_size} bytes)")
    return OUT / "DCN.so"


if __name__ == "__main__":
    sys.exit(0 if build(verbose="-v" in sys.argv) is not None or (OUT / "DCN.so").is_file() else 1)
