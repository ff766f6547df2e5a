#!/bin/bash
set -u
export PYTHONUNBUFFERED=1
echo "== full size, bench model, fp32 encoders"; timeout 300 python tools/debug_tiled.py 2240 3360 bench 2>&1 | tail -8
echo "== full size, bench model, bf16 encoders"; timeout 300 python tools/debug_tiled.py 2240 3360 bench bf16enc 2>&1 | tail -8
echo "== 448x672 test model bf16"; timeout 300 python tools/debug_tiled.py 448 672 bf16enc 2>&1 | tail -8
