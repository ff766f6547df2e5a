#!/bin/bash
set -u
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
echo "== world 2, overlap encoder"; timeout 300 $TR tools/tiled_check.py --height 2240 --width 3360 --iters 3 2>&1 | tail -1
echo "== world 1"; timeout 300 python tools/tiled_check.py --height 2240 --width 3360 --iters 3 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_tiled.py -m gpu -q -rP -p no:cacheprovider 2>&1 | tail -6
