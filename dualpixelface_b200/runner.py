"""Minimal runner standing in for pytorch_lightning (not installable here: no network).

Honours exactly the hooks the reference's main.py / model classes rely on (main.py:15-64,
src/model/model_selector.py:8-60 of the reference): ``train_dataloader / test_dataloader``, ``training_step``,
``test_step``, ``configure_optimizers``, ``self.log``, checkpoints named ``checkpoint_epoch=NN.ckpt`` holding
``{'state_dict': ...}``, and the 3-level JSON configuration (config_/config_manager.py:17-95).
"""
from __future__ import annotations

import json
from pathlib import Path
from runpy import run_path

import torch
import torch.nn as nn


class LightningModule(nn.Module):
    """The subset of pl.LightningModule the reference's models use."""

    def __init__(self):
        super().__init__()
        self.logged = {}
        self.current_epoch = 0
        self.global_step = 0

    def save_hyperparameters(self, *a, **k):
        pass

    def log(self, key, value, **kw):
        self.logged[key] = float(value.detach()) if torch.is_tensor(value) else float(value)


class Namespace:
    """Attribute view of a nested dict (the reference's config ``obj``, config_/config_manager.py:7-14)."""

    def __init__(self, d):
        for k, v in d.items():
            if isinstance(v, dict):
                v = Namespace(v)
            elif isinstance(v, (list, tuple)):
                v = [Namespace(x) if isinstance(x, dict) else x for x in v]
            setattr(self, k, v)

    def __contains__(self, k):
        return hasattr(self, k)


def load_config(config: str, workspace: str = "default", load_model=None, root: Path | str = ".", make_dirs: bool = True):
    """config_/<config>.json + src/model/<model>/<model_config>.json + dataloader/<dataset>/<cfg>.json (if present)."""
    root = Path(root)
    path = root / "config_" / f"{config}.json"
    if not path.is_file():
        raise FileNotFoundError(f"invalid config name: {path}")
    data = {"model": {}, "dataset": {}}
    data.update(json.loads(path.read_text()))
    data["load_model"] = str(Path(load_model).absolute()) if load_model else None
    data["sync_batch"] = data.get("accelerator") == "ddp"
    mcfg = root / "src" / "model" / data["model_name"] / f"{data['model_config']}.json"
    if not mcfg.is_file():
        raise FileNotFoundError(f"invalid model config: {mcfg}")
    data["model"] = json.loads(mcfg.read_text())
    dcfg = root / "dataloader" / data.get("dataset_name", "FaceDP") / f"{data.get('dataset_config', 'config')}.json"
    data["dataset"] = json.loads(dcfg.read_text()) if dcfg.is_file() else {"flip_lr": True, "dp_conversion": "given"}
    ws = root / "workspace" / data["model_name"] / workspace
    if make_dirs:
        (ws / "log").mkdir(parents=True, exist_ok=True)
        (ws / "output").mkdir(parents=True, exist_ok=True)
    data.update(model_path=str(ws.parent), workspace_path=str(ws), logger_path=str(ws / "log"), output_path=str(ws / "output"))
    return Namespace(data)


def model_selector(option, root: Path | str = "."):
    """src/model/<name>/mainmodel.py must define class <NAME upper> (model_selector.py:8-28 of the reference)."""
    ns = run_path(str(Path(root) / "src" / "model" / option.model_name / "mainmodel.py"))
    model = ns[option.model_name.upper()](option)
    if option.load_model is not None and option.mode != "train":
        ckpt = torch.load(option.load_model, map_location="cpu")
        sd = ckpt.get("state_dict", ckpt.get("model"))
        if sd is None:
            raise NotImplementedError("wrong checkpoint")
        model.load_state_dict(sd, strict=option.load_strict)
    return model


def optimizer_selector(params, option):
    if option.optim == "adam":
        return torch.optim.Adam(params, lr=float(option.init_lr), betas=(0.9, 0.999), eps=1e-5)
    if option.optim == "sgd":
        return torch.optim.SGD(params, lr=float(option.init_lr), momentum=0.9, weight_decay=2e-4)
    if option.optim == "rmsprop":
        return torch.optim.RMSprop(params, lr=float(option.init_lr), eps=1e-5)
    raise NotImplementedError("optimizer is not defined, please check your optimizer configuration !")


def scheduler_selector(optimizer, option):
    if option.scheduler == "steplr":
        return torch.optim.lr_scheduler.StepLR(optimizer, 35, 0.5)
    if option.scheduler == "explr":
        return torch.optim.lr_scheduler.ExponentialLR(optimizer, 0.5)
    if option.scheduler == "cosanneal":
        return torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, 500, 1e-6)
    if option.scheduler == "none":
        return None
    raise NotImplementedError("scheduler is not defined, please check your scheduler configuration !")


def _to_device(batch, device):
    return {k: (v.to(device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in batch.items()}


class Trainer:
    """fit / test loops over the model's own dataloaders; one process per GPU (DDP handled in parallel.py)."""

    def __init__(self, max_epochs=1, device="cuda", workspace_path=None, grad_sync=None):
        self.max_epochs, self.device, self.workspace_path, self.grad_sync = max_epochs, device, workspace_path, grad_sync

    def test(self, model, verbose=True):
        model.to(self.device).eval()
        outs = []
        with torch.no_grad():
            for i, batch in enumerate(model.test_dataloader()):
                outs.append(model.test_step(_to_device(batch, self.device), i))
        if hasattr(model, "test_epoch_end"):
            model.test_epoch_end(outs)
        return outs

    def fit(self, model):
        model.to(self.device).train()
        opts, scheds = model.configure_optimizers()
        for epoch in range(self.max_epochs):
            model.current_epoch = epoch
            for i, batch in enumerate(model.train_dataloader()):
                out = model.training_step(_to_device(batch, self.device), i)
                opts[0].zero_grad(set_to_none=True)
                out["loss"].backward()
                if self.grad_sync is not None:
                    self.grad_sync(model)
                opts[0].step()
                model.global_step += 1
                logs = ", ".join(f"{k} {float(v.detach()):.4f}" for k, v in out.get("log", {}).items())
                print(f"epoch {epoch} step {model.global_step}: loss {float(out['loss'].detach()):.4f}" + (f" ({logs})" if logs else ""))
            for s in scheds:
                s.step()
            if self.workspace_path:
                torch.save({"state_dict": model.state_dict(), "epoch": epoch},
                           Path(self.workspace_path) / f"checkpoint_epoch={epoch:02d}.ckpt")
