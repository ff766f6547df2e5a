"""Eager vs CUDA-graph replay of the untiled StereoDPNet inference forward at the bench shape."""
import sys
sys.path.insert(0, ".")
import torch
import bench
from dualpixelface_b200.synthetic import synthetic_batch
dev = torch.device("cuda", 0)
model = bench.build_model(dev)
batch = {k: v.to(dev) for k, v in synthetic_batch(4, 1120, 1680, seed=0).items()}
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
with torch.no_grad():
    ms_e = timeit(lambda: model(batch))
    want = model(batch)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2): model(batch)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = model(batch)
    ms_g = timeit(g.replay)
    g.replay(); torch.cuda.synchronize()
    same = torch.equal(out["pred_depth"], want["pred_depth"]) and torch.equal(out["pred_normal"], want["pred_normal"])
print(f"eager {ms_e:.3f} ms/step ({4e3 / ms_e:.1f} pairs/s), graph replay {ms_g:.3f} ms/step ({4e3 / ms_g:.1f} pairs/s), bit-identical: {same}")
