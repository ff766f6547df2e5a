"""Entry point with the reference's CLI (main.py:15-64 of the reference):
    python main.py --config <name> --workspace <name> [--load_model <ckpt>]
``mode`` in the config selects train / test ("demo" = a test config with batch 1).  Data is synthetic (the FaceDP
dataset is licensed and not present); under torchrun one process drives one GPU (NCCL)."""
import argparse
import os

import torch

from dualpixelface_b200.runner import Trainer, load_config, model_selector


def main():
    ap = argparse.ArgumentParser(description="Dual-Pixel Face Reconstruction on the sm_100a hot path")
    ap.add_argument("--config", type=str, required=True, help="config to run")
    ap.add_argument("--workspace", type=str, required=True, help="workspace name")
    ap.add_argument("--load_model", type=str, help="model path to load")
    a = ap.parse_args()
    opt = load_config(a.config, a.workspace, a.load_model)
    torch.manual_seed(1)
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    model = model_selector(opt)
    grad_sync = None
    if int(os.environ.get("WORLD_SIZE", 1)) > 1:
        from dualpixelface_b200.parallel import init_distributed, make_grad_sync
        from dualpixelface_b200.train_ops import set_sync_bn
        init_distributed()
        model.to(f"cuda:{local_rank}")
        if opt.sync_batch and opt.mode == "train":             # accelerator "ddp" => Trainer(sync_batchnorm=True) in the reference
            set_sync_bn(True)                                  # 3-D path: partial sums all-reduced inside the BN Functions
            model.feature_extraction = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model.feature_extraction)
        grad_sync = make_grad_sync(model)                      # hooks launch the bucket all-reduces during backward
    runner = Trainer(max_epochs=opt.epoch, device=f"cuda:{local_rank}",
                     workspace_path=opt.workspace_path if opt.mode == "train" else None, grad_sync=grad_sync)
    if opt.mode == "train":
        runner.fit(model)
    elif opt.mode == "test":
        outs = runner.test(model)
        print(f"tested {len(outs)} batches; pred_depth {tuple(outs[0]['pred_depth'].shape)}")
    else:
        raise NotImplementedError("Wrong mode !!")


if __name__ == "__main__":
    main()
