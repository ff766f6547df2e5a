#!/usr/bin/env python
"""NNet eval forward, one warm-up pass + N passes (for an ncu launch list: tools/summarize_ncu.py)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from dualpixelface_b200.runner import load_config, model_selector  # noqa: E402
from dualpixelface_b200.synthetic import synthetic_batch  # noqa: E402

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    model = model_selector(load_config("eval_faceDP_nnet", "test", root=ROOT, make_dirs=False), root=ROOT).cuda().eval()
    batch = {k: v.cuda() for k, v in synthetic_batch(4, 1120, 1680, seed=0).items()}
    with torch.no_grad():
        for _ in range(1 + n):
            model(batch)
    torch.cuda.synchronize()
