#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python tools/debug_tiled.py 192 160 2>&1 | tail -12
timeout 900 python -m pytest tests/test_gpu_tiled.py tests/test_gpu_teacher_forced.py tests/test_gpu_training.py tests/test_gpu_training_sdp.py "tests/test_gpu_kernels.py::test_fused_losses" -m gpu -q -rP -p no:cacheprovider > gpurun_out/r02d_pytest.log 2>&1
echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^E  |tiled \(|worst|ANM teacher|d\(out3|grad normal" gpurun_out/r02d_pytest.log | head -60
