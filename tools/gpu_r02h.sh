#!/bin/bash
set -u
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
echo "== world 1 graph"; timeout 300 python tools/tiled_check.py --height 2240 --width 3360 --iters 3 --graph 2>&1 | tail -1
echo "== world 2 graph"; timeout 300 $TR tools/tiled_check.py --height 2240 --width 3360 --iters 3 --graph 2>&1 | tail -1
