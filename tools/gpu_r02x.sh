#!/bin/bash
# round-2 GPU call X: final state -- full suite, full bench line, launch list, refreshed captures of the kernels that changed
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q -rP -p no:cacheprovider > gpurun_out/r02x_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02x_pytest.log
grep -E "passed|failed|error" gpurun_out/r02x_pytest.log | tail -3
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r02x_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02x_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "stage_ms", "e2e", "gpu_launches", "clocks")})
print("roofline", d["roofline"])
for k, v in sorted(d["roofline_kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:14]: print(f"  {k:28s} x{v['launches_per_step']:3d} {v['ms_per_step']:.3f} ms  {v['achieved']:8.1f} {v['unit']}  {v['frac']:.3f}")
print([(x["kernel"][:20], round(x["frac"], 4)) for x in d["roofline_extra"]])
for k in ("train", "psmnet", "config5"): print(k, json.dumps(d.get(k))[:500])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 5200 --csv --log-file gpurun_out/r02x_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu > gpurun_out/r02x_bench_under_ncu.log 2>&1
echo "launch list rc=$?"
cap() { local n=$1 k=$2 s=$3 c=$4; shift 4
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -o gpurun_out/r02x_$n -f "$@" > /dev/null 2>&1; echo "ncu $n rc=$?"; }
cap dcn3d dcn3d_kernel 2 1 python tools/bench_dcn.py --iters 2
cap regress regress_fwd 1 1 python tools/bench_membound.py
cap conv_t2 conv3d_tc_kernel 4 1 python tools/bench_conv_strided.py --only conv6 --iters 4
cap conv_kdfused_32x32 conv3d_kdfused 2 1 python tools/bench_conv.py --b 4 --cin 32 --cout 32
cap conv_kdfused_64x32 conv3d_kdfused 2 1 python tools/bench_conv.py --b 4 --cin 64 --cout 32
cap conv2d_96x32 conv2d_tc 2 1 python tools/bench_conv_res.py
ls gpurun_out | grep r02x
