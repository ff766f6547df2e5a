#!/bin/bash
# round-2 GPU call B: new kernels (conv2d_tc, gwc, deterministic issue) -- tests, A/B timings, bench
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q -rP -p no:cacheprovider -x --deselect tests/test_gpu_configs.py::test_config3_stereodpnet_training_step_1120x1680 > gpurun_out/r02b_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02b_pytest.log
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02b_pytest.log | head -30
{
for iss in 1 4; do
  echo "== DPF_CONV_ISSUERS=$iss"
  DPF_CONV_ISSUERS=$iss timeout 120 python tools/bench_conv.py --b 4 --cin 32 --cout 32
  DPF_CONV_ISSUERS=$iss timeout 120 python tools/bench_conv.py --b 4 --cin 64 --cout 32
  DPF_CONV_ISSUERS=$iss timeout 120 python tools/bench_conv.py --b 4 --cin 32 --cout 1
  DPF_CONV_ISSUERS=$iss timeout 120 python tools/bench_conv_strided.py
done
timeout 300 python tools/bench_conv2d.py
timeout 300 python tools/bench_membound.py
} > gpurun_out/r02b_micro.log 2>&1
cat gpurun_out/r02b_micro.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
echo "bench rc=$?"; tail -c 800 gpurun_out/r02b_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02b_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "stage_ms")}, d["e2e"]["value"])
print("roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 4))
for k, v in sorted(d["roofline_kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"]): print(f"  {k:28s} x{v['launches_per_step']:3d} {v['ms_per_step']:.3f} ms  {v['achieved']:8.1f} {v['unit']}  {v['frac']:.3f}")
for k in ("train", "psmnet", "costvol", "gpu_eager_oracle", "cpu_baseline"): print(k, json.dumps(d.get(k))[:900])
PY
DPF_CONV_ISSUERS=4 timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu > gpurun_out/r02b_bench_iss4.json 2>> gpurun_out/r02b_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r02b_bench_iss4.json').read().strip().splitlines()[-1]); print('ISSUERS=4:', d['value'], d['ms_per_step'], d['stage_ms'])"
