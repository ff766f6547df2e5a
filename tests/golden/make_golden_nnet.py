#!/usr/bin/env python
"""Golden fixtures for NNet, produced by running the UNMODIFIED reference
(src/model/nnet/mainmodel.py, through ref_shim) on seeded synthetic weights and inputs:

  state_keys_nnet.json   state_dict key -> shape of the reference's NNET
  model_nnet.npz         train-mode outputs (every 2nd pixel) + losses + a few parameter gradients, and eval outputs with BatchNorm statistics
                         calibrated by one oracle train pass (as make_golden.py does); prob_depth is kept at every 8th pixel

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_nnet.py
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
SIZE = (2, 256, 256)            # the SPP encoder's 64 x 64 average pool needs H/4, W/4 >= 64; train-mode BatchNorm needs B >= 2 there
GRAD_KEYS = ("classify.2.weight", "dres0.0.0.weight", "dres3.2.0.weight", "convs.0.0.weight", "convs.6.0.weight",
             "normal_module.wc0.0.0.weight", "normal_module.pool2.0.0.weight", "normal_module.n_convs.6.0.weight",
             "feature_extraction.firstconv.0.0.weight", "feature_extraction.lastconv.2.weight")

if __name__ == "__main__":
    assert ref_shim.reference_available(), "needs the reference checkout at /root/reference"
    torch.manual_seed(1)
    m = ref_shim.build_reference_model("nnet")
    shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
    (HERE / "state_keys_nnet.json").write_text(json.dumps(shapes, indent=0))
    st = synth_state({k: tuple(v) for k, v in shapes.items()}, seed=1)
    batch = synthetic_batch(*SIZE, training=True, seed=0)
    out = {}
    # ---- train mode: forward + losses + gradients of the reference itself ------------------------------------------------
    m.load_state_dict(st, strict=False)
    m.train()
    with ref_shim._in_reference_tree():
        res = m(dict(batch))
        res["final_loss"].backward()
    out["train/pred_depth_s2"] = res["pred_depth"][..., ::2, ::2].detach().float().numpy()      # every 2nd pixel: fixture size
    out["train/pred_normal_s2"] = res["pred_normal"][..., ::2, ::2].detach().float().numpy()
    for k in ("smoothL1_loss", "cosine_loss", "final_loss"):
        out[f"train/{k}"] = np.float32(res[k].item())
    params = dict(m.named_parameters())
    for k in GRAD_KEYS:
        out[f"train/grad/{k}"] = params[k].grad.detach().float().numpy()
    print("train losses", {k: float(out[f"train/{k}"]) for k in ("smoothL1_loss", "cosine_loss", "final_loss")})
    # ---- eval mode with calibrated running statistics -------------------------------------------------------------------
    stats = {}
    with torch.no_grad():
        O.nnet_forward(dict(batch), st, True, stats=stats)
    st = O.calibrate_running_stats(st, stats)
    m.load_state_dict(st, strict=False)
    m.eval()
    m.normal_module.grid_check = False                       # the pixel grid is registered lazily, once (normal_module_.py:57-65)
    m.normal_module._parameters.pop("grid", None)
    with ref_shim._in_reference_tree(), torch.no_grad():
        res = m(dict(batch))
    for k in ("pred_depth", "pred_normal", "ref_feature"):
        out[f"eval/{k}"] = res[k].float().numpy()
    out["eval/prob_depth_s8"] = res["prob_depth"][..., ::8, ::8].float().numpy()
    print("eval pred_depth range", float(res["pred_depth"].min()), float(res["pred_depth"].max()))
    np.savez_compressed(HERE / "model_nnet.npz", **out)
    print("model_nnet.npz", (HERE / "model_nnet.npz").stat().st_size // 1024, "KiB")
