#!/usr/bin/env python
"""Golden fixtures for StereoNet, produced by running the UNMODIFIED reference
(src/model/stereonet/mainmodel.py, through ref_shim) on seeded synthetic weights and inputs:

  state_keys_stereonet.json   state_dict key -> shape of the reference's STEREONET
  model_stereonet.npz         eval outputs (BatchNorm statistics calibrated by one oracle train pass, as make_golden.py does) and
                              train-mode outputs + loss + a few parameter gradients

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_stereonet.py
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE))

import ref_shim  # noqa: E402
from dualpixelface_b200.synthetic import synth_state, synthetic_batch  # noqa: E402
from oracle import dpf_oracle as O  # noqa: E402

torch.set_num_threads(8)
GRAD_KEYS = ("conv3d_alone.weight", "filter.0.0.0.weight", "feature_extraction.conv_alone.weight",
             "edge_aware_refinements.0.conv2d_out.weight", "feature_extraction.downsample.0.weight")

if __name__ == "__main__":
    assert ref_shim.reference_available(), "needs the reference checkout at /root/reference"
    torch.manual_seed(1)
    m = ref_shim.build_reference_model("stereonet")
    shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
    (HERE / "state_keys_stereonet.json").write_text(json.dumps(shapes, indent=0))
    st = synth_state({k: tuple(v) for k, v in shapes.items()}, seed=1)
    batch = synthetic_batch(2, 64, 96, training=True, seed=0)
    out = {}
    # ---- train mode: forward + loss + gradients of the reference itself --------------------------------------------------
    m.load_state_dict(st, strict=False)
    m.train()
    with ref_shim._in_reference_tree():
        res = m(dict(batch))
        res["final_loss"].backward()
    out["train/pred_depth"] = res["pred_depth"].detach().float().numpy()
    out["train/final_loss"] = np.float32(res["final_loss"].item())
    params = dict(m.named_parameters())
    for k in GRAD_KEYS:
        out[f"train/grad/{k}"] = params[k].grad.detach().float().numpy()
    # ---- eval mode with calibrated running statistics -------------------------------------------------------------------
    stats = {}
    with torch.no_grad():
        O.stereonet_forward(dict(batch), st, True, stats=stats)
    st = O.calibrate_running_stats(st, stats)
    m.load_state_dict(st, strict=False)
    m.eval()
    with ref_shim._in_reference_tree(), torch.no_grad():
        res = m(dict(batch))
    for k in ("pred_depth", "prob_depth", "ref_feature"):
        out[f"eval/{k}"] = res[k].float().numpy()
    print("eval pred_depth range", float(res["pred_depth"].min()), float(res["pred_depth"].max()))
    np.savez_compressed(HERE / "model_stereonet.npz", **out)
    print("model_stereonet.npz", (HERE / "model_stereonet.npz").stat().st_size // 1024, "KiB")
