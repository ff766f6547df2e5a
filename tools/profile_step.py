"""torch.profiler kernel table for one bench step (cheap alternative to an ncu launch list)."""
import sys
sys.path.insert(0, ".")
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from dualpixelface_b200.synthetic import synthetic_batch
dev = torch.device("cuda", 0)
model = bench.build_model(dev)
batch = {k: v.to(dev) for k, v in synthetic_batch(bench.B, bench.H, bench.W, seed=0).items()}
with torch.no_grad():
    for _ in range(3):
        model(batch)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(2):
            model(batch)
        torch.cuda.synchronize()
evs = [e for e in prof.key_averages() if e.device_time_total > 0]
evs.sort(key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in evs)
print(f"total device time per step: {tot / 2 / 1000:.2f} ms")
for e in evs[:45]:
    print(f"{e.device_time_total / 2 / 1000:8.3f} ms  {e.count // 2:4d}x  {e.key[:110]}")
