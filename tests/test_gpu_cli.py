"""GPU: the reference's entry contract end to end -- config_/<name>.json -> model_selector -> Trainer.test / Trainer.fit -- for
every model directory under src/model (main.py:15-64 of the reference drives exactly these calls)."""
import math

import pytest
import torch

from dualpixelface_b200.runner import Trainer, load_config, model_selector

from conftest import ROOT

pytestmark = pytest.mark.gpu

EVAL = {"eval_faceDP": (128, 160), "eval_faceDP_psmnet": (256, 256), "eval_faceDP_stereonet": (128, 160), "eval_faceDP_nnet": (256, 256)}
TRAIN = {"train_faceDP_psmnet": (256, 256), "train_faceDP_stereonet": (128, 160), "train_faceDP_nnet": (256, 256)}


def _option(cfg, size):
    opt = load_config(cfg, "pytest", root=ROOT, make_dirs=False)
    opt.synthetic_size, opt.synthetic_batches, opt.batch_size = size, 2, 2
    return opt


@pytest.mark.parametrize("cfg", sorted(EVAL))
def test_trainer_test_loop(cfg):
    opt = _option(cfg, EVAL[cfg])
    assert opt.mode == "test"
    torch.manual_seed(1)
    outs = Trainer(device="cuda").test(model_selector(opt, root=ROOT), verbose=False)
    assert len(outs) == 2
    h, w = EVAL[cfg]
    for o in outs:
        assert o["pred_depth"].shape[0] == 2 and o["pred_depth"].shape[-2:] == (h, w) and o["pred_depth"].is_cuda
        assert o["ref_feature"].shape[0] == 2
        if o.get("pred_normal") is not None:
            assert o["pred_normal"].shape == (2, 1, 3, h, w)


@pytest.mark.parametrize("cfg", sorted(TRAIN))
def test_trainer_fit_one_epoch(cfg, capsys):
    """Two optimizer steps on synthetic pairs: finite losses, parameters of the 3-D trunk and of the encoder move."""
    opt = _option(cfg, TRAIN[cfg])
    assert opt.mode == "train"
    torch.manual_seed(1)
    model = model_selector(opt, root=ROOT)
    before = {k: v.detach().clone() for k, v in model.named_parameters() if v.requires_grad}
    Trainer(max_epochs=1, device="cuda").fit(model)
    assert model.global_step == 2
    losses = [float(line.split("loss ")[1].split()[0]) for line in capsys.readouterr().out.splitlines() if line.startswith("epoch")]
    assert len(losses) == 2 and all(math.isfinite(v) for v in losses)
    moved = [k for k, v in model.named_parameters() if k in before and not torch.equal(v.detach().cpu(), before[k].cpu())]
    assert any("feature_extraction" in k for k in moved) and any("feature_extraction" not in k for k in moved)
    # StereoNet's BasicBlock.conv2 is constructed but never applied (src/model/stereonet/modules.py:23): no gradient, as in the reference
    expected = [k for k in before if not (cfg.endswith("stereonet") and ".conv2." in k)]
    assert len(moved) > 0.9 * len(expected), (len(moved), len(expected))
