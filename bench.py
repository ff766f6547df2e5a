#!/usr/bin/env python
"""Benchmark of the DualPixelFace stereo hot path on B200 (contract: see the build brief / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): StereoDPNet inference on synthetic 1120x1680 dual-pixel pairs, batch 4 per GPU,
bf16 hot path.  One step = one forward pass of the model (encoder -> cost volume -> 3-D aggregation -> regression ->
normal branch) over one batch.  `value` = pairs/s with inputs resident in HBM; `e2e` = the same through the model's
public call with pinned host buffers (H2D of the images, D2H of depth + normals inside the timed region).
For N > 1 launch with torchrun: independent image pairs per rank (weak scaling), no data-path collective in inference.
`--impl reference` times the reference algorithm on the host CPU (the oracle port; the reference has no CPU path of its
own and cannot be installed here), on a bounded sample of the same workload.

Extra blocks on the same JSON line (each guarded: a failure is reported inside its block and never loses the contract line):
  "train"    BASELINE config 3: StereoDPNet fwd + bwd (depth + normal losses) + Adam step at 1120x1680, batch 8 per GPU (or the
             largest of 8/4/2 that fits, stated); under torchrun every rank runs it with the overlapping bucketed NCCL gradient
             all-reduce (parallel.GradSync), so the 1 -> 8 GPU curve of the driver contains the collective
  "psmnet"   BASELINE config 1 (1 x 448x448 eval) and config 4 (512x768 crops: eval and fwd+bwd+optimizer, also under torchrun)
  "costvol"  north_star kernel (1): dpf_costvol_fwd (concat / diff / gwc) alone, CUDA events, GB/s against the HBM roofline
  "config5"  BASELINE config 5: one 2240x3360 pair -- at N = 1 the untiled forward, under torchrun row tiles over the N GPUs with
             halo exchange (dualpixelface_b200/tiled.py); ms per pair (max over ranks), halo bytes and exchanges per pair
  "gpu_eager_oracle"  the oracle's own PyTorch code on the SAME B200 (fp32 eager, cuDNN), as context for the speed-up
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

H, W, B = 1120, 1680, 4
WORKLOAD = f"stereodpnet_infer_{H}x{W}_b{B}"
AGG_FLOP_PER_VOXEL = 644544            # SURVEY.md 8a-5, forward, per quarter-res voxel (D*H4*W4 voxels per pair)
D3D_OFFSET_SCALE = tuple(float(v) for v in os.environ.get("DPF_BENCH_D3D_OFFSET_SCALE", "0.0078125,0.01").split(","))
VOL_BYTES_PER_QPIX = 1152              # 128 B read + 1024 B written per quarter-res pixel (concat volume)


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.is_file():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], tf=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured (MEASURED_PEAKS.json, sustained)")
    return dict(hbm=6650.0, tf=1400.0, src="fallback (B200_PROFILING.md)")


def state_shapes(name):
    return {k: tuple(v) for k, v in json.loads((ROOT / "tests" / "golden" / f"state_keys_{name}.json").read_text()).items()}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def build_model(device, config="eval_faceDP", name="stereodpnet"):
    from dualpixelface_b200.runner import load_config, model_selector
    from dualpixelface_b200.synthetic import synth_state
    opt = load_config(config, "bench", root=ROOT, make_dirs=False)
    model = model_selector(opt, root=ROOT)
    model.load_state_dict(synth_state(state_shapes(name), seed=1), strict=False)
    # The two offset convolutions of the deformable layers are scaled so that the sampling offsets are about one voxel (sigma
    # 0.9 / 1.0; a trained deformable conv's regime -- the class is zero-initialised in the reference).  With the unscaled
    # random weights the offsets have sigma 115 / 8 voxels: 98 % / 75 % of the samples then fall outside the 4-plane volume and
    # are skipped by the kernel, and a row tile (config 5) would need its whole neighbour as halo (tools/dcn_offset_stats.py).
    ne = getattr(model, "normal_estimator", None)
    if ne is not None and getattr(ne, "use_deform", False):
        with torch.no_grad():
            for layer, sc in ((ne.deform_conv1, D3D_OFFSET_SCALE[0]), (ne.deform_conv2, D3D_OFFSET_SCALE[1])):
                layer.conv_offset.weight.mul_(sc)
                layer.conv_offset.bias.mul_(sc)
        model.refresh()
    return model.to(device).eval()


def train_block(model_name, b, h, w, steps, warmup, dev, rank, world):
    """fwd + bwd + optimizer step of `model_name` at b x h x w per GPU; returns the metrics dict (max over ranks)."""
    from dualpixelface_b200 import ops
    from dualpixelface_b200.runner import optimizer_selector
    from dualpixelface_b200.synthetic import synthetic_batch
    cfg = "train_faceDP" if model_name == "stereodpnet" else "train_faceDP_psmnet"
    model = build_model(dev, cfg, model_name).train()
    opt = optimizer_selector(model.parameters(), model.option)
    sync = None
    if world > 1:
        from dualpixelface_b200.parallel import make_grad_sync
        sync = make_grad_sync(model)
    batch = {k: v.to(dev) for k, v in synthetic_batch(b, h, w, training=True, seed=1000 + rank).items()}

    def step():
        opt.zero_grad(set_to_none=True)
        res = model(batch)
        res["final_loss"].backward()
        if sync is not None:
            sync()
        opt.step()
        return res

    for _ in range(warmup):
        res = step()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    n0 = ops.launch_count()
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        res = step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = t.item()
    out = {"workload": f"{model_name}_train_{h}x{w}_b{b}", "batch_per_gpu": b, "steps": steps, "warmup": warmup,
           "ms_per_step": ms / steps, "pairs_per_s": world * b * steps / (ms * 1e-3), "n_gpus": world,
           "what": "forward + backward + optimizer step (Adam); depth" + (" + normal" if getattr(model, "predict_normal", False) else "") + " losses",
           "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "gpu_launches": int(ops.launch_count() - n0),
           "final_loss": float(res["final_loss"].detach()),
           "aggregation_tflop_per_step_fwd_bwd": 3 * AGG_FLOP_PER_VOXEL * 8 * (h // 4) * (w // 4) * b / 1e12}
    if sync is not None:
        out["grad_allreduce"] = {"bytes": int(sum(f.numel() * 4 for f in sync.flat)), "buckets": len(sync.buckets),
                                 "launched_during_backward_per_step": sync.launched_in_backward / (steps + warmup), "backend": "nccl"}
        sync.remove()
    del model, opt, batch, res
    torch.cuda.empty_cache()
    return out


def guarded_train_block(model_name, batches, h, w, steps, warmup, dev, rank, world):
    """Largest per-GPU batch of `batches` that fits; errors are returned inside the block instead of raised."""
    err = None
    for b in batches:
        try:
            return train_block(model_name, b, h, w, steps, warmup, dev, rank, world)
        except torch.cuda.OutOfMemoryError as e:
            err = f"batch {b}: out of memory ({str(e)[:80]})"
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            return {"error": f"{type(e).__name__}: {str(e)[:300]}"}
    return {"error": err}


def run_train(args):
    """`--mode train`: only the training block, as its own JSON line (`--model stereodpnet|psmnet --batch B --height H --width W`)."""
    from dualpixelface_b200 import ops
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    ops.lib()
    blk = train_block(args.model, args.batch, args.height, args.width, args.steps, max(args.warmup, 3), dev, rank, world)
    if rank == 0:
        print(json.dumps({"metric": f"{args.model} training DP-pairs/sec (fwd+bwd+optimizer)", "value": blk["pairs_per_s"],
                          "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                          "ms_per_step": blk["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "bf16", "data": "synthetic", "config": {"workload": blk["workload"], "parallelism": f"dp{world}"},
                          **{k: v for k, v in blk.items() if k not in ("pairs_per_s", "ms_per_step", "workload", "steps", "warmup")}}))
    if world > 1:
        torch.distributed.destroy_process_group()


def _time_ms(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def costvol_block(dev, peaks):
    """north_star kernel (1), dpf_costvol_fwd, timed alone: integer-shift concat / difference / group-wise-correlation volumes in
    bf16 NDHWC.  Algorithmic bytes = both feature maps read once + the volume written once.  The config-2-sized
    case moves 542 MB per launch (>> the 126 MB L2); the config-4 / config-1 cases are listed as they are (L2-resident inputs)."""
    from dualpixelface_b200 import ops
    shifts = [-1, 0, 0, 0, 1, 1, 2, 2]                       # int(costrange), psmnet/modules.py:229
    out = []
    for name, (b, h4, w4) in (("c2-sized 4x280x420", (4, 280, 420)), ("config 4: 8x128x192", (8, 128, 192)), ("config 1: 1x112x112", (1, 112, 112))):
        g = torch.Generator(device=dev).manual_seed(0)
        ref = torch.randn(b, h4, w4, 32, device=dev, generator=g).to(torch.bfloat16)
        tgt = torch.randn(b, h4, w4, 32, device=dev, generator=g).to(torch.bfloat16)
        for mode, groups in (("concat", 0), ("diff", 0), ("gwc", 8)):
            cv = ops.costvol_channels(mode, 32, groups)
            nbytes = b * h4 * w4 * (2 * 32 * 2 + 8 * cv * 2)
            ms = _time_ms(lambda: ops.costvol_fwd(ref, tgt, shifts, mode, groups), 20, 3)
            gbs = nbytes / (ms * 1e-3) / 1e9
            out.append({"shape": name, "mode": mode, "ms": round(ms, 4), "bytes": nbytes, "achieved": round(gbs, 1), "unit": "GB/s",
                        "frac": round(gbs / peaks["hbm"], 4), "frac_of_8TBs_nominal": round(gbs / 8000.0, 4)})
    return {"kernel": "costvol_fwd_kernel (dpf_costvol_fwd)", "bound": "hbm", "peak": peaks["hbm"], "cases": out}


def psmnet_block(dev, rank, world, steps):
    """BASELINE config 1 (PSMNet eval, one 448x448 pair) and config 4 (512x768 crops: eval forward and training step)."""
    from dualpixelface_b200.synthetic import synthetic_batch
    out = {}
    try:
        model = build_model(dev, "eval_faceDP_psmnet", "psmnet")
        with torch.no_grad():
            for key, (b, h, w) in (("config1_eval_1x448x448", (1, 448, 448)), ("config4_eval_8x512x768", (8, 512, 768))):
                batch = {k: v.to(dev) for k, v in synthetic_batch(b, h, w, seed=rank).items()}
                ms = _time_ms(lambda: model(batch), max(steps, 5), 3)
                out[key] = {"ms_per_step": round(ms, 3), "pairs_per_s_per_gpu": round(b / (ms * 1e-3), 1)}
        del model
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out["eval_error"] = f"{type(e).__name__}: {str(e)[:300]}"
    out["config4_train"] = guarded_train_block("psmnet", (8, 4, 2), 512, 768, max(3, min(steps, 5)), 3, dev, rank, world)
    return out


def other_models_block(dev, steps):
    """the two other cost-volume models of the reference on the same kernels -- eval forward of 4 x 1120x1680 pairs
    on this GPU (synthetic 'calibrated' weights), with the time of each stage."""
    from dualpixelface_b200.synthetic import synthetic_batch
    out = {}
    for name in ("nnet", "stereonet"):
        model = batch = None
        try:
            model = build_model(dev, f"eval_faceDP_{name}", name)
            batch = {k: v.to(dev) for k, v in synthetic_batch(4, H, W, seed=0).items()}
            with torch.no_grad():
                ms = _time_ms(lambda: model(batch), max(steps, 5), 3)
                model.stage_events = []
                model(batch)
                torch.cuda.synchronize()
                ev = model.stage_events
                model.stage_events = None
            out[name] = {"ms_per_step": round(ms, 3), "pairs_per_s": round(4 / (ms * 1e-3), 1), "shape": [4, H, W],
                         "stage_ms": {n: round(a.elapsed_time(b_), 3) for (_, a), (n, b_) in zip(ev[:-1], ev[1:])}}
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
        del model, batch
        torch.cuda.empty_cache()
    return out


def gpu_eager_oracle_block(dev):
    """Context line: the oracle's own PyTorch code (= the reference's algorithm, restated) executed on the SAME B200 in fp32
    eager mode with cuDNN (TF32 off), StereoDPNet eval forward of ONE 1120x1680 pair.  Not a baseline to beat by itself (it is
    unoptimised eager code), but it separates 'GPU vs CPU' from 'this implementation vs a plain GPU port'."""
    from dualpixelface_b200.synthetic import synth_state, synthetic_batch
    from oracle import dpf_oracle as O                      # checker code, timed as a stated baseline only
    tf = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        st = {k: v.to(dev) for k, v in synth_state(state_shapes("stereodpnet"), seed=1).items()}
        batch = {k: v.to(dev) for k, v in synthetic_batch(1, H, W, seed=0).items()}
        with torch.no_grad():
            ms = _time_ms(lambda: O.stereodpnet_forward(dict(batch), st, False), 3, 1)
        return {"ms_per_pair": round(ms, 2), "pairs_per_s": round(1e3 / ms, 3), "dtype": "f32", "what": "oracle/dpf_oracle.py "
                f"stereodpnet_forward, eval, 1 x {H}x{W}, torch eager + cuDNN on this GPU, TF32 off"}
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {str(e)[:300]}"}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf
        torch.cuda.empty_cache()


def config5_block(dev, rank, world, steps):
    """BASELINE config 5: ONE 2240x3360 pair; under torchrun split into row tiles over the N GPUs with halo exchange
    (dualpixelface_b200/tiled.py), else the untiled single-GPU forward.  Device time of one pair (max over ranks)."""
    from dualpixelface_b200.synthetic import synthetic_batch
    h, w = 2240, 3360
    try:
        model = build_model(dev)
        batch = {k: v.to(dev) for k, v in synthetic_batch(1, h, w, seed=0).items()}
        out = {"workload": f"stereodpnet_infer_{h}x{w}_b1", "n_gpus": world}
        with torch.no_grad():
            if world == 1:
                ms = _time_ms(lambda: model(batch), max(steps, 3), 2)
                out.update(ms_per_pair=round(ms, 3), mode="untiled, one GPU")
            else:
                from dualpixelface_b200.tiled import TiledStereoDPNet
                tm = TiledStereoDPNet(model, h, rank, world)
                tm(batch); tm(batch)
                tm.t.bytes_exchanged = tm.t.exchanges = 0
                tm.t.log = []
                tm(batch)
                torch.cuda.synchronize()
                sent, nex = tm.t.bytes_exchanged, tm.t.exchanges
                by_layer = {}                                  # halo traffic of rank 0 per exchanged tensor class
                for shape, dt, rows, nbytes in tm.t.log:
                    k = f"{'x'.join(map(str, shape))} {dt}, rows {rows[0]}+{rows[1]}"
                    e = by_layer.setdefault(k, [0, 0])
                    e[0] += 1
                    e[1] += nbytes
                tm.t.log = None
                torch.distributed.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = max(steps, 3)
                e0.record()
                for _ in range(n):
                    tm(batch)
                e1.record()
                torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
                out.update(ms_per_pair=round(t.item(), 3), mode=f"row tiles over {world} GPUs, per-layer halo exchange (NCCL p2p) in the 3-D path, "
                           "overlap-recompute encoder", rows_rank0=list(tm.t.tiles[0]), halo_bytes_sent_rank0=int(sent), exchanges_per_pair=int(nex),
                           d3d_halo_rows=tm._hd, d3d_reach_ok=bool(tm.check_reach()),
                           halo_by_tensor_rank0={k: {"exchanges": v[0], "bytes_sent": v[1]} for k, v in by_layer.items()})
                # the tiled forward is launch-bound (35 exchanges + ~500 kernels for a few ms of GPU work per rank): the same pass
                # captured in a CUDA graph, NCCL halo exchanges included (tiled.TiledStereoDPNet.capture)
                try:
                    replay, _static, _sout = tm.capture(batch)
                    replay(); replay()
                    torch.cuda.synchronize()
                    torch.distributed.barrier()
                    e0.record()
                    for _ in range(n):
                        replay()
                    e1.record()
                    torch.cuda.synchronize()
                    tg = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
                    torch.distributed.all_reduce(tg, op=torch.distributed.ReduceOp.MAX)
                    out["ms_per_pair_cuda_graph"] = round(tg.item(), 3)
                    del replay, _static, _sout
                except Exception as e:  # noqa: BLE001
                    out["cuda_graph_error"] = f"{type(e).__name__}: {str(e)[:200]}"
        del model, batch
        torch.cuda.empty_cache()
        return out
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {str(e)[:300]}"}


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from this round's `ncu --set full` captures, by kernel key
    (profiles/r02_traffic.json, written by tools/summarize_ncu.py --traffic)."""
    p = ROOT / "profiles" / "r02_traffic.json"
    return json.loads(p.read_text()) if p.is_file() else {}


def cpu_reference_pairs_per_s(steps, warmup, sample_hw=(448, 672)):
    """Reference algorithm (oracle port) on the host CPU: `steps` forwards of ONE pair at sample_hw, scaled to
    full-resolution pair equivalents by pixel count (the network is fully convolutional)."""
    from dualpixelface_b200.synthetic import synth_state, synthetic_batch
    from oracle import dpf_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    st = synth_state(state_shapes("stereodpnet"), seed=1)
    batch = synthetic_batch(1, sample_hw[0], sample_hw[1], seed=0)
    with torch.no_grad():
        for _ in range(warmup):
            O.stereodpnet_forward(dict(batch), st, False)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.stereodpnet_forward(dict(batch), st, False)
        dt = time.perf_counter() - t0
    frac = (sample_hw[0] * sample_hw[1]) / float(H * W)
    return steps * frac / dt, dt / steps, cores, f"{steps} x 1 pair at {sample_hw[0]}x{sample_hw[1]} ({frac:.3f} of a {H}x{W} pair), fp32, eval"


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 10))
    val, sec, cores, sample = cpu_reference_pairs_per_s(steps, max(1, min(args.warmup, 1)))
    print(json.dumps({
        "impl": "reference", "metric": "StereoDPNet DP-pairs/sec", "value": val, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "batch_per_gpu": B, "height": H, "width": W},
        "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "extrapolated: a 448x672 pair scaled by pixel count to 1120x1680-pair equivalents; the reference has no CPU "
                                 "path of its own (its deformable conv is CUDA-only) and cannot be installed here, so the oracle port is timed"},
        "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_ours(args):
    from dualpixelface_b200 import ops
    from dualpixelface_b200.synthetic import synthetic_batch
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    ops.lib()
    model = build_model(dev, "eval_faceDP" if args.model == "stereodpnet" else "eval_faceDP_psmnet", args.model)
    host = synthetic_batch(B, H, W, seed=rank)
    batch = {k: v.to(dev) for k, v in host.items()}

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            out = model(batch)
        # ---------------- device-resident throughput -----------------------------------------------------
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        barrier()
        n0 = ops.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = model(batch)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = ops.launch_count() - n0
        clocks = sampler.stop() if rank == 0 else None
        # ---------------- per-stage device times (CUDA events on the launching stream) --------------------
        stage_ms = {}
        reps = min(args.steps, 5)
        for _ in range(reps):
            model.stage_events = []
            model(batch)
            torch.cuda.synchronize()
            ev = model.stage_events
            for (_, a), (name, b_) in zip(ev[:-1], ev[1:]):
                stage_ms[name] = stage_ms.get(name, 0.0) + a.elapsed_time(b_) / reps
        model.stage_events = None
        # ---------------- per-launch device times of the dominant kernels (events around single launches) -
        ops.KERNEL_TIMING = {}
        for _ in range(reps):
            model(batch)
        torch.cuda.synchronize()
        kernels = {}
        for key, recs in ops.KERNEL_TIMING.items():
            t_ms = sum(a.elapsed_time(b_) for a, b_, _, _ in recs)
            work = sum(w_ for _, _, w_, _ in recs)
            kernels[key] = {"launches_per_step": len(recs) // reps, "ms_per_step": t_ms / reps, "avg_launch_ms": t_ms / len(recs),
                            "unit": "TFLOP/s" if recs[0][3] == "flop" else "GB/s",
                            "achieved": work / (t_ms * 1e-3) / (1e12 if recs[0][3] == "flop" else 1e9),
                            "work_per_launch": work / len(recs)}
        ops.KERNEL_TIMING = None
        # ---------------- end to end: pinned host buffers -> H2D -> forward -> D2H ------------------------
        # host buffers: images travel as bf16 (the encoder's input precision: its first op casts to bf16 anyway), results come
        # back as fp16 (disparity in [-4,12] px: 2^-8 px resolution; normals in [-1,1]) -- 2x fewer host bytes in each direction
        # than fp32, which is what limited the 8-GPU end-to-end scaling (all ranks share one NUMA node's copy bandwidth)
        pin = {k: (v.to(torch.bfloat16) if k in ("left", "right") else v).pin_memory() for k, v in host.items()}
        res_d = torch.empty(B, 1, H, W, dtype=torch.float16).pin_memory()
        res_n = torch.empty(B, 1, 3, H, W, dtype=torch.float16).pin_memory()
        h2d = sum(v.numel() * v.element_size() for v in pin.values())
        d2h = res_d.numel() * 2 + (res_n.numel() * 2 if getattr(model, "predict_normal", False) else 0)

        # Double-buffered: the H2D copy of step i+1 (copy stream) and the D2H read of step i-1 (second copy stream) overlap the
        # forward of step i; every step's copies are enqueued and completed inside the timed region.
        cur = torch.cuda.current_stream()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        dbuf = [{k: torch.empty_like(v, device=dev) for k, v in pin.items()} for _ in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        ev_done, ev_read = torch.cuda.Event(), torch.cuda.Event()

        def upload(i):
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_free[i % 2])                  # the forward that last read this buffer has finished
                for k, v in pin.items():
                    dbuf[i % 2][k].copy_(v, non_blocking=True)
                ev_in[i % 2].record(s_in)

        def e2e_run(n):
            for e in ev_free:
                e.record(cur)
            ev_read.record(s_out)
            upload(0)
            for i in range(n):
                if i + 1 < n:
                    upload(i + 1)
                cur.wait_event(ev_in[i % 2])
                o = model(dbuf[i % 2])
                ev_free[i % 2].record(cur)
                ev_done.record(cur)
                od = o["pred_depth"].to(torch.float16)
                on = o["pred_normal"].to(torch.float16) if o["pred_normal"] is not None else None
                ev_done.record(cur)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_done)
                    res_d.copy_(od, non_blocking=True)
                    if on is not None:
                        res_n.copy_(on, non_blocking=True)
                    ev_read.record(s_out)
                od.record_stream(s_out)
                if on is not None:
                    on.record_stream(s_out)
            cur.wait_event(ev_read)                              # the last result is on the host before the region ends

        e2e_run(2)
        barrier()
        e0.record()
        e2e_run(args.steps)
        e1.record()
        barrier()
        ms_e2e = e0.elapsed_time(e1)

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    # ---------------- extra blocks: every rank takes part in the ones that contain a collective ----------------
    del model, batch, dbuf, out
    torch.cuda.empty_cache()
    extras = {}
    if not args.no_extras and args.model == "stereodpnet" and (H, W) == (1120, 1680):
        tsteps = max(3, min(args.steps, 5))
        extras["train"] = guarded_train_block("stereodpnet", (8, 4, 2), H, W, tsteps, 3, dev, rank, world)
        extras["psmnet"] = psmnet_block(dev, rank, world, tsteps)
        extras["config5"] = config5_block(dev, rank, world, tsteps)
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    peaks = measured_peaks()
    h4, w4 = H // 4, W // 4
    agg_flops = AGG_FLOP_PER_VOXEL * 8 * h4 * w4 * B
    agg_tf = agg_flops / (stage_ms["aggregation"] * 1e-3) / 1e12
    vol_gbs = VOL_BYTES_PER_QPIX * h4 * w4 * B / (stage_ms["cost_volume"] * 1e-3) / 1e9
    reg_bytes = (32 * h4 * w4 + 4 * H * W) * B
    for v in kernels.values():
        v["frac"] = v["achieved"] / (peaks["tf"] if v["unit"] == "TFLOP/s" else peaks["hbm"])
    # headline roofline = the DOMINANT kernel of the step by device time (not the best-looking one)
    dom_key, dom = max(kernels.items(), key=lambda kv: kv[1]["ms_per_step"])
    traffic = measured_traffic().get(dom_key)
    line = {
        "metric": "StereoDPNet DP-pairs/sec" if args.model == "stereodpnet" else "PSMNet DP-pairs/sec (extra, not the contract metric)",
        "value": world * B * args.steps / (ms * 1e-3), "unit": "pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "height": H, "width": W, "parallelism": f"dp{world}",
                   "l2": "per-step working set (~6 GB of activations) is far larger than the 126 MB L2; no explicit flush",
                   "weights": "seeded synthetic ('calibrated' style), eval-mode BatchNorm folded into the conv epilogues; the D3D offset "
                              f"convolutions scaled by {D3D_OFFSET_SCALE[0]:g} / {D3D_OFFSET_SCALE[1]:g} (offsets of about one voxel)",
                   "e2e_io": "images host->device as bf16, disparity + normals device->host as fp16 (pinned, double-buffered)"},
        "e2e": {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "stage_ms": {k: round(v, 3) for k, v in stage_ms.items()},
        # achieved = ALGORITHMIC work per launch / average launch duration (CUDA events around the single launches, inside the
        # step); for the D3D layers the work is the true 35 -> 64 / 64 -> 64 FLOPs, not the zero-padded 64 -> 64 the kernel runs
        "roofline": {"kernel": dom_key, "why": f"dominant kernel of the step by device time ({dom['ms_per_step']:.2f} of {ms / args.steps:.2f} ms)",
                     "bound": "tensor" if dom["unit"] == "TFLOP/s" else "hbm",
                     "achieved": dom["achieved"], "peak": peaks["tf"] if dom["unit"] == "TFLOP/s" else peaks["hbm"], "unit": dom["unit"],
                     "frac": dom["frac"], "traffic": traffic["bytes_per_launch"] if traffic else None,
                     "traffic_source": traffic["source"] if traffic else None, "peak_source": peaks["src"],
                     "work_per_launch": dom["work_per_launch"], "avg_launch_ms": dom["avg_launch_ms"],
                     "launches_per_step": dom["launches_per_step"]},
        "roofline_kernels": {k: {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items()} for k, v in kernels.items()},
        "roofline_extra": [
            {"kernel": "3-D aggregation stage (all launches of the conv engine)", "bound": "tensor", "achieved": agg_tf,
             "peak": peaks["tf"], "unit": "TFLOP/s", "frac": agg_tf / peaks["tf"], "flops_per_step": agg_flops},
            {"kernel": "cost volume stage (StereoDPNet ASM: sample + mask convs + statistics + blend)", "bound": "hbm", "achieved": vol_gbs,
             "peak": peaks["hbm"], "unit": "GB/s", "frac": vol_gbs / peaks["hbm"], "bytes_per_step": VOL_BYTES_PER_QPIX * h4 * w4 * B},
            {"kernel": "regression stage (regress_fwd_kernel)", "bound": "hbm", "achieved": reg_bytes / (stage_ms["regression"] * 1e-3) / 1e9,
             "peak": peaks["hbm"], "unit": "GB/s", "frac": reg_bytes / (stage_ms["regression"] * 1e-3) / 1e9 / peaks["hbm"],
             "bytes_per_step": reg_bytes}],
    }
    line.update(extras)
    if world == 1 and not args.no_extras:
        try:
            line["costvol"] = costvol_block(dev, peaks)
        except Exception as e:  # noqa: BLE001
            line["costvol"] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
        line["gpu_eager_oracle"] = gpu_eager_oracle_block(dev)
        if args.model == "stereodpnet" and (H, W) == (1120, 1680):
            line["other_models"] = other_models_block(dev, max(3, min(args.steps, 5)))
    if world == 1 and not args.no_cpu:
        val, sec, cores, sample = cpu_reference_pairs_per_s(3, 1)
        line["cpu_baseline"] = {"value": val, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample,
                                "note": "extrapolated: a 448x672 pair scaled by pixel count to 1120x1680-pair equivalents"}
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="only the contract measurements (profiling runs)")
    ap.add_argument("--mode", default="infer", choices=["infer", "train"], help="train = extra fwd+bwd+optimizer line")
    ap.add_argument("--model", default="stereodpnet", choices=["stereodpnet", "psmnet"])
    ap.add_argument("--batch", type=int, default=None, help="pairs per GPU (default: 4 inference = BASELINE configs[1], 8 training)")
    ap.add_argument("--height", type=int, default=H)
    ap.add_argument("--width", type=int, default=W)
    a = ap.parse_args()
    if a.mode == "train":
        a.batch = a.batch or 8
    else:                                  # other inference shapes (e.g. BASELINE config 5: --batch 1 --height 2240 --width 3360)
        H, W, B = a.height, a.width, a.batch or B
        WORKLOAD = f"{a.model}_infer_{H}x{W}_b{B}"
    if a.impl == "reference":
        run_reference(a)
    elif a.mode == "train":
        run_train(a)
    else:
        run_ours(a)
