#!/usr/bin/env python
"""Golden fixtures for the NON-DEFAULT configurations of the reference, produced by running the UNMODIFIED reference
(through ref_shim) with the corresponding config overrides:

  sdp_nodeform_sampling / sdp_nodeform_nosampling   STEREODPNET with use_deform=false and use_sampling=true / false
        (src/model/stereodpnet/normal_module.py:45-56,159-163,181-183): whole-model eval outputs
  psm_gwcnet8                                        PSMNET with cost_volume='gwcnet', group_num=8 (psmnet/modules.py:243-271)

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_variants.py  ->  variants.npz
(use_deform=true needs the compiled CUDA op and therefore has no CPU fixture; its kernels are pinned on the GPU against
oracle/_ref/DCN.so by tests/test_gpu_dcn_reference.py.)
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


def run(name, overrides, cfg, hw, out, tag):
    m = ref_shim.build_reference_model(name, **overrides)
    shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
    out[f"{tag}/state_keys"] = np.frombuffer(json.dumps(shapes).encode(), dtype=np.uint8)
    st = synth_state({k: tuple(v) for k, v in shapes.items()}, seed=1)
    batch = synthetic_batch(2, hw[0], hw[1], training=True, seed=0)
    fwd = O.psmnet_forward if name == "psmnet" else O.stereodpnet_forward
    stats = {}
    with torch.no_grad():
        fwd(dict(batch), st, True, cfg=cfg, stats=stats)          # calibrated BN statistics (oracle train pass, as make_golden.py)
    st = O.calibrate_running_stats(st, stats)
    m.load_state_dict(st, strict=False)
    m.eval()
    if name == "stereodpnet":
        ref_shim.reset_shift_cache(m)
    with ref_shim._in_reference_tree(), torch.no_grad():
        res = m(dict(batch))
    out[f"{tag}/pred_depth"] = res["pred_depth"].float().numpy()
    if res.get("pred_normal") is not None:
        out[f"{tag}/pred_normal"] = res["pred_normal"].float().numpy()
    print(tag, "pred_depth range", float(res["pred_depth"].min()), float(res["pred_depth"].max()))


if __name__ == "__main__":
    assert ref_shim.reference_available(), "needs the reference checkout at /root/reference"
    torch.manual_seed(1)
    out = {}
    for samp in (True, False):
        tag = f"sdp_nodeform_{'sampling' if samp else 'nosampling'}"
        run("stereodpnet", dict(use_deform=False, use_sampling=samp), dict(O.SDP_CFG, use_deform=False, use_sampling=samp), (64, 96), out, tag)
    run("psmnet", dict(cost_volume="gwcnet", group_num=8), dict(O.PSM_CFG, cost_volume="gwcnet", group_num=8), (256, 256), out, "psm_gwcnet8")
    np.savez_compressed(HERE / "variants.npz", **out)
    print("variants.npz", (HERE / "variants.npz").stat().st_size // 1024, "KiB")
