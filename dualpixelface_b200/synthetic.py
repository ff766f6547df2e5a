"""Seeded synthetic inputs and weights.

There is no dataset and no checkpoint in this environment, so parity tests, the smoke run and the
bench all use: unit-variance noise images, a pin-hole ``K`` with f = 4000 px, the camera-1
``abvalue`` of the FaceDP reader (dataloader/FaceDP/path_reader.py:26 in the reference) and random
targets.  Weights are generated *per state-dict key* from a hash of the key, so the reference model
(golden generation), the oracle and the CUDA product all get bit-identical parameters without a
weight file having to travel.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Mapping, Sequence

import torch

ABVALUE_CAM1 = (32.98, -26996.49)


def synthetic_batch(batch: int, height: int, width: int, training: bool = False, seed: int = 0,
                    device: str | torch.device = "cpu") -> Dict[str, torch.Tensor]:
    """A FaceDP-shaped batch dict (keys follow dataloader/FaceDP/loader.py:149-155 of the reference)."""
    g = torch.Generator().manual_seed(seed)
    out = {
        "left": torch.randn(batch, 3, height, width, generator=g),
        "right": torch.randn(batch, 3, height, width, generator=g),
        "K": torch.tensor([[4000.0, 0.0, width / 2.0], [0.0, 4000.0, height / 2.0], [0.0, 0.0, 1.0]]).repeat(batch, 1, 1),
        "abvalue": torch.tensor([ABVALUE_CAM1]).repeat(batch, 1),
    }
    if training:
        out["disp"] = torch.rand(batch, height, width, generator=g) * 8.0 - 2.0
        out["mask"] = torch.ones(batch, height, width)
        out["depth"] = torch.rand(batch, height, width, generator=g) * 500.0 + 800.0
        out["idepth"] = torch.rand(batch, height, width, generator=g)
        out["normal"] = torch.randn(batch, 3, height, width, generator=g)
    return {k: v.to(device) for k, v in out.items()}


def _key_generator(key: str, seed: int) -> torch.Generator:
    return torch.Generator().manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def synth_tensor(key: str, shape: Sequence[int], seed: int = 1, style: str = "calibrated") -> torch.Tensor:
    """One parameter / buffer, determined only by (key, shape, seed, style).

    style 'ref_init'   : the reference's own init distributions (N(0, sqrt(2/(k*C_out))) for convs,
                         BN weight 1 / bias 0, running stats 0 / 1; src/model/stereodpnet/mainmodel.py:51-65).
    style 'calibrated' : fan-in scaled convs and mildly perturbed norm parameters / running statistics so
                         that eval-mode activations stay O(1) through the 28-layer aggregation (with
                         'ref_init' + eval BN the soft-argmin saturates, SURVEY.md section 8c caveats).
    """
    shape = tuple(int(s) for s in shape)
    # MaskingAttention registers its InstanceNorm twice (self.normalize and mask_convs.3.1 are the same
    # module, src/module/asm/asm.py:138-146 of the reference): both names must get the same values
    key = key.replace("attention_layer.normalize.", "attention_layer.mask_convs.3.1.")
    g = _key_generator(key, seed)
    leaf = key.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.long)
    if leaf == "running_mean":
        return torch.zeros(shape) if style == "ref_init" else torch.randn(shape, generator=g) * 0.05
    if leaf == "running_var":
        return torch.ones(shape) if style == "ref_init" else torch.rand(shape, generator=g) * 0.4 + 0.8
    if key.endswith("costrange") or key.endswith(".grid"):
        raise KeyError(f"{key} is a derived constant, not a synthetic weight")
    if len(shape) == 1:
        if leaf == "bias":
            if style == "ref_init":
                return torch.zeros(shape)
            return torch.randn(shape, generator=g) * 0.05
        if leaf == "weight":                       # BN / IN gamma or PReLU slope
            if shape == (1,):                      # nn.PReLU(init=0.05) slopes
                return torch.full(shape, 0.05)
            if style == "ref_init":
                return torch.ones(shape)
            return torch.rand(shape, generator=g) * 0.4 + 0.8
    if len(shape) >= 3:                            # conv / deconv / deformable-conv kernels
        ksz = 1
        for s in shape[2:]:
            ksz *= s
        if style == "ref_init":
            std = math.sqrt(2.0 / (ksz * shape[0]))
        else:
            fan_in = ksz * shape[1]
            std = math.sqrt(2.0 / fan_in)
            if "conv_offset" in key:
                std *= 0.25                          # keep deformable offsets sub-voxel
            if shape[0] == 1 or (len(shape) == 5 and shape[0] <= 3):
                std *= 0.5                           # regression / normal heads: keep logits O(1)
        return torch.randn(shape, generator=g) * std
    return torch.randn(shape, generator=g) * 0.05


def synth_state(shapes: Mapping[str, Sequence[int]], seed: int = 1, style: str = "calibrated") -> Dict[str, torch.Tensor]:
    """State dict for every key in ``shapes`` (derived constants such as ANM's costrange/grid are skipped)."""
    out = {}
    for key, shape in shapes.items():
        if key.endswith("costrange") or key.endswith(".grid"):
            continue
        out[key] = synth_tensor(key, shape, seed, style)
    return out
