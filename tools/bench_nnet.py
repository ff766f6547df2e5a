#!/usr/bin/env python
"""NNet on one B200: eval time per stage at the StereoDPNet bench shape (B x 1120 x 1680), and one training step.  Prints one JSON object."""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from dualpixelface_b200.runner import load_config, model_selector  # noqa: E402
from dualpixelface_b200.synthetic import synthetic_batch  # noqa: E402


def timed(fn, warm, reps):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--size", type=int, nargs=2, default=(1120, 1680))
    ap.add_argument("--train-batch", type=int, default=2)
    args = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    H, W = args.size
    model = model_selector(load_config("eval_faceDP_nnet", "test", root=ROOT, make_dirs=False), root=ROOT).cuda().eval()
    batch = {k: v.cuda() for k, v in synthetic_batch(args.batch, H, W, training=True, seed=0).items()}
    out = {"model": "nnet", "shape": [args.batch, H, W]}
    with torch.no_grad():
        ms = timed(lambda: model(batch), 3, 10)
        out["eval_ms_per_step"] = ms
        out["eval_pairs_per_s"] = args.batch / ms * 1e3
        stage = {}
        for _ in range(5):
            model.stage_events = []
            model(batch)
            torch.cuda.synchronize()
            ev = model.stage_events
            for (_, a), (name, b_) in zip(ev[:-1], ev[1:]):
                stage[name] = stage.get(name, 0.0) + a.elapsed_time(b_) / 5
        model.stage_events = None
        out["eval_stage_ms"] = {k: round(v, 3) for k, v in stage.items()}
    del batch
    torch.cuda.empty_cache()
    model.train()
    tb = {k: v.cuda() for k, v in synthetic_batch(args.train_batch, H, W, training=True, seed=1).items()}
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)

    def step():
        opt.zero_grad(set_to_none=True)
        model(tb)["final_loss"].backward()
        opt.step()

    t0 = time.time()
    ms_t = timed(step, 2, 3)
    out["train_ms_per_step"] = ms_t
    out["train_pairs_per_s"] = args.train_batch / ms_t * 1e3
    out["train_batch"] = args.train_batch
    out["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
    print(json.dumps(out))
