"""BASELINE config 5 on N GPUs (torchrun): one high-resolution pair, row tiles with halo exchange (dualpixelface_b200/tiled.py).
Checks the gathered tiled result against the untiled single-GPU model on rank 0 and times both.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/tiled_check.py [--height 2240 --width 3360]
"""
import argparse, json, os, sys
sys.path.insert(0, ".")
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--height", type=int, default=2240); ap.add_argument("--width", type=int, default=3360)
ap.add_argument("--iters", type=int, default=5); ap.add_argument("--fp32-encoder", action="store_true")
ap.add_argument("--graph", action="store_true", help="also capture the tiled forward (incl. NCCL halo exchanges) in a CUDA graph")
ap.add_argument("--weights", default="calibrated", choices=["calibrated", "bench"],
                help="calibrated: BatchNorm statistics from an oracle pass (numerically sane outputs, for the error check); bench: bench.py's seeded weights")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from bench import build_model
from dualpixelface_b200 import ops
from dualpixelface_b200.synthetic import synthetic_batch
from dualpixelface_b200.tiled import TiledStereoDPNet
ops.lib()
model = build_model(dev)
if a.weights == "calibrated":
    sys.path.insert(0, "tests")
    from test_gpu_models import calibrated_state          # test infrastructure (uses the oracle for the calibration pass)
    st, _ = calibrated_state("stereodpnet", synthetic_batch(2, 128, 160, training=True, seed=0))
    model.load_state_dict(st, strict=False)
    model.to(dev).eval()
if a.fp32_encoder:
    model.encoder_autocast = False
batch = {k: v.to(dev) for k, v in synthetic_batch(1, a.height, a.width, seed=0).items()}
tm = TiledStereoDPNet(model, a.height, rank, world)

def timed(fn, iters):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()

res = tm.gather(tm(batch))
tm.t.bytes_exchanged = 0; tm.t.exchanges = 0
tm.stage_events = []
one = tm(batch)
torch.cuda.synchronize()
halo_bytes, exchanges = tm.t.bytes_exchanged, tm.t.exchanges
ev = tm.stage_events
stage_ms = {b_[0]: round(a_[1].elapsed_time(b_[1]), 3) for a_, b_ in zip(ev[:-1], ev[1:])}
tm.stage_events = None
ms_tiled = timed(lambda: tm(batch), a.iters)
ms_graph, graph_err = None, None
if a.graph:
    try:
        replay, sbatch, sout = tm.capture(batch)
        replay(); torch.cuda.synchronize()
        graph_err = float((sout["pred_depth"] - one["pred_depth"]).abs().max())
        ms_graph = timed(replay, a.iters)
    except Exception as e:  # noqa: BLE001
        graph_err = f"{type(e).__name__}: {str(e)[:200]}"
out = {"world": world, "height": a.height, "width": a.width, "ms_tiled": round(ms_tiled, 3), "halo_bytes_sent_per_rank": halo_bytes,
       "exchanges_per_pass": exchanges, "rows": list(tm.t.tiles[rank]), "stage_ms_rank0": stage_ms, "ms_tiled_cuda_graph": ms_graph,
       "graph_vs_eager_max_diff": graph_err, "d3d_halo_rows": tm._hd, "reach_ok": tm.check_reach()}
if rank == 0:
    with torch.no_grad():
        ref = model(batch)
        torch.cuda.synchronize()
    d = (res["pred_depth"] - ref["pred_depth"]).abs()
    n = (res["pred_normal"] - ref["pred_normal"]).abs()
    out.update(disp_max_err=round(d.max().item(), 5), disp_mean_err=round(d.mean().item(), 6), normal_max_err=round(n.max().item(), 5),
               normal_mean_err=round(n.mean().item(), 6))
if world > 1:
    dist.barrier()
if rank == 0:
    # untiled timing on rank 0 alone (the other ranks idle)
    with torch.no_grad():
        for _ in range(2): model(batch)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters): model(batch)
        e1.record(); torch.cuda.synchronize()
        out["ms_untiled_1gpu"] = round(e0.elapsed_time(e1) / a.iters, 3)
    out["speedup"] = round(out["ms_untiled_1gpu"] / ms_tiled, 3)
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
