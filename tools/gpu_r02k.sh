#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_configs.py tests/test_gpu_models.py tests/test_gpu_training.py tests/test_gpu_conv2d_tc.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -12
{
timeout 120 python tools/bench_conv.py --b 4 --cin 32 --cout 32; timeout 120 python tools/bench_conv.py --b 4 --cin 64 --cout 32
timeout 120 python tools/bench_conv_strided.py
} 2>&1 | tee gpurun_out/r02k_micro.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu > gpurun_out/r02k_bench.json 2>gpurun_out/r02k_bench.err; tail -c 300 gpurun_out/r02k_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02k_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "stage_ms")})
for k, v in sorted(d["roofline_kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:10]: print(f"  {k:28s} x{v['launches_per_step']:3d} {v['ms_per_step']:.3f} ms  {v['achieved']:8.1f} {v['unit']}  {v['frac']:.3f}")
print([round(x["frac"], 4) for x in d["roofline_extra"]])
PY
