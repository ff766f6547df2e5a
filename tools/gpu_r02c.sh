#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_tiled.py tests/test_gpu_teacher_forced.py tests/test_gpu_configs.py tests/test_gpu_training.py tests/test_gpu_training_sdp.py -m gpu -q -rP -p no:cacheprovider > gpurun_out/r02c_pytest.log 2>&1
echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^E  |tiled|teacher|worst|grad " gpurun_out/r02c_pytest.log | head -80
timeout 300 python tools/tiled_check.py --height 2240 --width 3360 --iters 3 2>&1 | tail -3
