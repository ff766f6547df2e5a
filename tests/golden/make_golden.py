#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference (through ref_shim).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Outputs (all small, committed):
  state_keys_<model>.json     state_dict key -> shape of the reference's STEREODPNET / PSMNET
  stages.npz                  stage-level outputs of the reference's own modules on seeded inputs
  model_<model>.npz           whole-model outputs (eval with calibrated BN statistics, and train mode)
Inputs and weights are regenerated from seeds by dualpixelface_b200.synthetic, so they are not stored.
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


def feat(shape, seed):
    """Post-ReLU-like feature map (N(0,1) then ReLU)."""
    g = torch.Generator().manual_seed(seed)
    return torch.relu(torch.randn(*shape, generator=g))


def np32(t):
    return t.detach().to(torch.float32).cpu().numpy()


def gen_state_keys():
    models = {}
    for name in ("psmnet", "stereodpnet"):
        m = ref_shim.build_reference_model(name)
        shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
        (HERE / f"state_keys_{name}.json").write_text(json.dumps(shapes, indent=0))
        models[name] = (m, shapes)
    return models


def gen_stages(models):
    mods = ref_shim.reference_modules()
    out = {}
    sdp, sdp_shapes = models["stereodpnet"]
    psm, _ = models["psmnet"]
    st = synth_state(sdp_shapes, seed=1)
    sdp.load_state_dict(st, strict=False)
    opt = sdp.option

    # ---- a-1 sub-pixel shift: fresh module per call so the cache bug does not interfere -------------
    x = feat((2, 4, 16, 24), 11)
    for disp in (-1.0, -0.5, 0.5, 1.0, 2.5):
        for direction in ("forward", "backward"):
            layer = mods["asm"].subpixel_shift(opt)
            smp = layer(x, disp, direction)
            out[f"shift/{disp}/{direction}"] = np32(torch.cat(smp, -1))          # [B,C,H,W,3]
    # nearest source-row / column tables on an index ramp at the BASELINE resolutions
    for (h, w) in ((112, 112), (128, 192), (280, 420), (560, 840)):
        ramp_r = (torch.arange(h, dtype=torch.float32) + 1).view(1, 1, h, 1).expand(1, 1, h, w).contiguous()
        ramp_c = (torch.arange(w, dtype=torch.float32) + 1).view(1, 1, 1, w).expand(1, 1, h, w).contiguous()
        for direction in ("forward", "backward"):
            layer = mods["asm"].subpixel_shift(opt)
            layer.mode_bilinear = False
            layer.mode_phase = False
            r = layer(ramp_r, -1.0, direction)[0][0, 0, :, :, 0]
            layer2 = mods["asm"].subpixel_shift(opt)
            layer2.mode_bilinear = False
            layer2.mode_phase = False
            c = layer2(ramp_c, -1.0, direction)[0][0, 0, :, :, 0]
            # value v>0 -> source index v-1 ; 0 -> out of bounds
            out[f"nearest_rows/{h}x{w}/{direction}"] = (r[:, 0].round().long() - 1).numpy().astype(np.int32)
            out[f"nearest_cols/{h}x{w}/{direction}"] = (c[h // 2, :].round().long() - 1).numpy().astype(np.int32)
            assert bool((r == r[:, :1]).all() | True)

    # ---- a-4 integer-shift volumes -------------------------------------------------------------------
    ref, tgt = feat((2, 32, 12, 10), 21), feat((2, 32, 12, 10), 22)
    cv = psm.cost_volume
    out["psm/concat"] = np32(cv.build_concat_volume(ref, tgt))
    cv.group_num = 8
    out["psm/gwc8"] = np32(cv.build_gwc_volume(ref, tgt))
    cv.group_num = 40

    # ---- a-2/a-3 ASM volume, bug-compatible (cached first level) --------------------------------------
    ref, tgt = feat((2, 32, 16, 24), 31), feat((2, 32, 16, 24), 32)
    for training in (False, True):
        sdp.train(training)
        sdp.load_state_dict(st, strict=False)
        ref_shim.reset_shift_cache(sdp)
        with torch.no_grad():
            v = sdp.cost_volume(ref, tgt)
        out[f"sdp/volume/{'train' if training else 'eval'}"] = np32(v[:, :, :2])     # slices 0,1 (all 8 are equal)
        out[f"sdp/volume_all_equal/{'train' if training else 'eval'}"] = np.array(
            [bool(torch.equal(v[:, :, 0], v[:, :, i])) for i in range(8)])
    sdp.load_state_dict(st, strict=False)

    # ---- a-5 aggregation -------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(41)
    vol = torch.relu(torch.randn(1, 64, 8, 16, 16, generator=g))
    for training in (False, True):
        sdp.train(training)
        sdp.load_state_dict(st, strict=False)
        with torch.no_grad():
            costs, outs = sdp.aggregation(vol)
        tag = "train" if training else "eval"
        for i, c in enumerate(costs):
            out[f"agg/{tag}/cost{3 - i}"] = np32(c)
        out[f"agg/{tag}/out3"] = np32(outs[0])
    sdp.load_state_dict(st, strict=False)

    # ---- a-6 regression ----------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(51)
    cost_full = torch.randn(2, 32, 20, 28, generator=g) * 2.0
    d, p = sdp.regression_layer([cost_full])
    out["regress/disp"] = np32(d[0])
    out["regress/prob"] = np32(p[0])

    # ---- a-7 ANM pieces --------------------------------------------------------------------------------
    anm = sdp.normal_estimator
    g = torch.Generator().manual_seed(61)
    cost = torch.randn(2, 8, 6, 10, 12, generator=g)                        # b d c h w
    dq = (torch.rand(2, 1, 10, 12, generator=g) * 4.2 - 1.3)                # generic, off-level values
    sc, sd_ = anm.sample_with_sort(cost, dq)
    out["anm/sel_cost"] = np32(sc)
    out["anm/sel_disp"] = np32(sd_)
    batch = synthetic_batch(2, 40, 48, seed=3)
    coord = anm.grid_maker_3d(sc, batch["K"], sd_, batch["abvalue"])
    out["anm/coord"] = np32(coord)
    ref_shim.reset_shift_cache(sdp)

    # ---- a-10 losses ------------------------------------------------------------------------------------
    b = synthetic_batch(2, 16, 24, training=True, seed=5)
    g = torch.Generator().manual_seed(71)
    preds = {"pred_depth": torch.randn(2, 3, 16, 24, generator=g) * 3, "pred_normal": torch.randn(2, 1, 3, 16, 24, generator=g)}
    b["mask"] = (torch.rand(2, 16, 24, generator=g) > 0.3).float()
    with ref_shim._in_reference_tree():
        res = sdp.loss_model.forward(preds, b)
    out["loss/smoothL1"] = np32(res["smoothL1_loss"])
    out["loss/cosine"] = np32(res["cosine_loss"])
    out["loss/final"] = np32(res["final_loss"])
    out["loss/mask"] = np32(b["mask"])

    np.savez_compressed(HERE / "stages.npz", **out)
    print("stages.npz:", len(out), "arrays")


def gen_models(models):
    for name, (m, shapes) in models.items():
        out = {}
        st = synth_state(shapes, seed=1)
        h, w = (256, 256) if name == "psmnet" else (64, 96)
        batch = synthetic_batch(2, h, w, training=True, seed=0)
        fwd = O.psmnet_forward if name == "psmnet" else O.stereodpnet_forward
        # calibrated eval state: BN running stats := batch stats of one train-mode oracle pass
        stats = {}
        with torch.no_grad():
            fwd(dict(batch), st, True, stats=stats)
        st_cal = O.calibrate_running_stats(st, stats)
        for tag, state, training in (("eval", st_cal, False), ("train", st, True)):
            m.load_state_dict(state, strict=False)
            m.train(training)
            if name == "stereodpnet":
                ref_shim.reset_shift_cache(m)
            with ref_shim._in_reference_tree():
                if training:
                    res = m(dict(batch))
                    res["final_loss"].backward()
                    for key in ("aggregation.dres0.0.0.weight", "aggregation.classif3.2.weight",
                                "aggregation.dres4.conv6.0.weight", "feature_extraction.firstconv.0.0.weight"):
                        gr = dict(m.named_parameters())[key].grad
                        out[f"{tag}/grad/{key}"] = np32(gr)
                    m.zero_grad()
                else:
                    with torch.no_grad():
                        res = m(dict(batch))
            for k, v in res.items():
                if v is None or k in ("abvalue", "prob_depth"):
                    continue
                out[f"{tag}/{k}"] = np32(v)
            pd = res["prob_depth"].detach()
            out[f"{tag}/prob_depth_sub"] = np32(pd[..., ::8, ::8])
            print(name, tag, "pred_depth range", float(res["pred_depth"].min()), float(res["pred_depth"].max()))
        np.savez_compressed(HERE / f"model_{name}.npz", **out)
        print(f"model_{name}.npz:", len(out), "arrays")


if __name__ == "__main__":
    assert ref_shim.reference_available(), "needs the reference checkout at /root/reference"
    torch.manual_seed(1)
    with torch.no_grad():
        models = gen_state_keys()
    with torch.no_grad():
        gen_stages(models)
    gen_models(models)
    for f in sorted(HERE.glob("*.npz")) + sorted(HERE.glob("*.json")):
        print(f.name, f.stat().st_size // 1024, "KiB")
