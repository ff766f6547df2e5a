#!/bin/bash
# round-2 GPU call T (final evidence pass): full gpu test-suite, full bench line, launch list, ncu --set full captures
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q -rP -p no:cacheprovider > gpurun_out/r02t_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02t_pytest.log
grep -E "passed|failed|error" gpurun_out/r02t_pytest.log | tail -5
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r02t_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02t_bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "stage_ms", "e2e", "gpu_launches", "clocks")})
    print("roofline", d["roofline"])
    for k, v in sorted(d["roofline_kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:16]: print(f"  {k:28s} x{v['launches_per_step']:3d} {v['ms_per_step']:.3f} ms  {v['achieved']:8.1f} {v['unit']}  {v['frac']:.3f}")
    print([(x["kernel"][:20], round(x["frac"], 4)) for x in d["roofline_extra"]])
    for k in ("train", "psmnet", "config5", "costvol", "gpu_eager_oracle", "cpu_baseline"): print(k, json.dumps(d.get(k))[:900])
except Exception as e: print("parse failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r02t_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu > gpurun_out/r02t_bench_under_ncu.log 2>&1
echo "launch list rc=$?"
cap() { # name regex skip count cmd...
  local n=$1 k=$2 s=$3 c=$4; shift 4
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -o gpurun_out/r02t_$n -f "$@" > /dev/null 2>&1
  echo "ncu $n rc=$?"
}
cap dcn3d dcn3d_kernel 2 2 python tools/bench_dcn.py --iters 2
cap costvol costvol_fwd 3 3 python tools/bench_membound.py
cap regress regress_fwd 1 1 python tools/bench_membound.py
cap conv_s2 conv3d_s2 2 2 python tools/bench_conv_strided.py
cap conv_kdfused_64x32 conv3d_kdfused 2 1 python tools/bench_conv.py --b 4 --cin 64 --cout 32
cap conv_kdfused_32x32 conv3d_kdfused 2 1 python tools/bench_conv.py --b 4 --cin 32 --cout 32
cap conv2d conv2d_tc 2 2 python tools/bench_conv2d.py
ls -la gpurun_out | tail -20
cap head conv3d_head 3 1 python tools/bench_head.py
cap conv_t2 "conv3d_tc_kernel<2" 2 1 python tools/bench_conv_strided.py --only conv6
python tools/bench_dcn.py > gpurun_out/r02t_dcn_inbounds.log 2>&1; python tools/bench_head.py >> gpurun_out/r02t_dcn_inbounds.log 2>&1; python tools/graph_check.py >> gpurun_out/r02t_dcn_inbounds.log 2>&1; cat gpurun_out/r02t_dcn_inbounds.log
