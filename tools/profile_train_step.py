"""torch.profiler kernel table for one TRAINING step (fwd + bwd + optimizer) of the bench's train mode."""
import argparse, sys
sys.path.insert(0, ".")
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from dualpixelface_b200.runner import optimizer_selector
from dualpixelface_b200.synthetic import synthetic_batch
ap = argparse.ArgumentParser()
ap.add_argument("--model", default="stereodpnet"); ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--height", type=int, default=1120); ap.add_argument("--width", type=int, default=1680)
a = ap.parse_args()
dev = torch.device("cuda", 0)
model = bench.build_model(dev, "train_faceDP" if a.model == "stereodpnet" else "train_faceDP_psmnet", a.model).train()
opt = optimizer_selector(model.parameters(), model.option)
batch = {k: v.to(dev) for k, v in synthetic_batch(a.batch, a.height, a.width, training=True, seed=0).items()}
def step():
    opt.zero_grad(set_to_none=True)
    res = model(batch)
    res["final_loss"].backward()
    opt.step()
for _ in range(3):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
step(); torch.cuda.synchronize()
print(f"wall per step {1e3 * (time.perf_counter() - t0):.1f} ms")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
evs = [e for e in prof.key_averages() if e.device_time_total > 0]
evs.sort(key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in evs)
print(f"total device time per step: {tot / 1000:.2f} ms")
for e in evs[:60]:
    print(f"{e.device_time_total / 1000:8.3f} ms  {e.count:4d}x  {e.key[:120]}")
