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
    # The reference guards this with `option.mode is not 'train'` (model_selector.py:18) -- an identity comparison with a
    # literal that is true for every config string read from JSON, so a given --load_model is loaded in EVERY mode
    # (fine-tuning / resuming in train mode included).  Same behaviour here, stated instead of accidental.
    if option.load_model is not None:
        sd = load_checkpoint_state(option.load_model)
        model.load_state_dict(sd, strict=option.load_strict)
    return model


class _AnyObject:
    """Stand-in for classes a Lightning checkpoint pickles next to the weights (the reference calls save_hyperparameters(),
    so 'hyper_parameters' holds its config object `config_.config_manager.obj`, not importable here)."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"state": state})


def load_checkpoint_state(path):
    """state_dict of a checkpoint file: the repo's own {'state_dict', 'epoch'} files, the reference's released
    pytorch_lightning checkpoints ({'state_dict', 'hyper_parameters': <pickled config object>, 'callbacks', ...}) and plain
    {'model': ...} files.  Tensors-only files load with the safe `weights_only=True`; a file that pickles foreign classes is
    re-read with an unpickler that maps every class it cannot import to an inert stand-in, and only 'state_dict' is kept."""
    import pickle
    try:
        ckpt = torch.load(path, map_location="cpu", weights_only=True)
    except pickle.UnpicklingError:
        class _Unpickler(pickle.Unpickler):
            def find_class(self, module, name):
                if module.split(".")[0] in ("torch", "collections", "numpy", "builtins", "_codecs"):
                    return super().find_class(module, name)
                return _AnyObject

        class _Pickle:
            Unpickler = _Unpickler
            __name__ = "pickle"
            load = staticmethod(lambda f, **kw: _Unpickler(f, **kw).load())

        ckpt = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_Pickle)
    sd = ckpt.get("state_dict", ckpt.get("model")) if isinstance(ckpt, dict) else None
    if sd is None:
        raise NotImplementedError("wrong checkpoint")
    return {k: v for k, v in sd.items() if torch.is_tensor(v)}


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

    @staticmethod
    def _rank_world():
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
        return 0, 1

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
        rank, world = self._rank_world()
        if world > 1:
            import torch.distributed as dist
            from .parallel import broadcast_module_state
            broadcast_module_state(model)                      # replicas start identical (what DDP does at construction)
            if hasattr(model, "data_seed_offset"):
                model.data_seed_offset = rank * 100003         # every rank draws its own synthetic pairs
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
            if self.workspace_path and rank == 0:              # one writer; the other ranks wait so nobody reads a partial file
                torch.save({"state_dict": model.state_dict(), "epoch": epoch},
                           Path(self.workspace_path) / f"checkpoint_epoch={epoch:02d}.ckpt")
            if world > 1:
                dist.barrier()
