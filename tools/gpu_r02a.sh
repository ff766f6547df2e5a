#!/bin/bash
# round-2 GPU call A: full gpu test-suite (with measured values printed), the bench line, launch list, ncu captures
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q -rP -p no:cacheprovider > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
timeout 300 python tools/bench_dcn.py --offc 96 > gpurun_out/r02a_dcn.log 2>&1
timeout 300 python tools/bench_dcn.py --offc 81 >> gpurun_out/r02a_dcn.log 2>&1
cat gpurun_out/r02a_dcn.log
timeout 300 python tools/bench_membound.py > gpurun_out/r02a_membound.log 2>&1
cat gpurun_out/r02a_membound.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r02a_bench.err; head -c 1500 gpurun_out/r02a_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02a_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu > gpurun_out/r02a_bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dcn3d_kernel -s 2 -c 2 -o gpurun_out/r02a_dcn3d -f \
    python tools/bench_dcn.py --iters 2 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:costvol_fwd -s 3 -c 3 -o gpurun_out/r02a_costvol -f \
    python tools/bench_membound.py > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:regress_fwd -s 1 -c 1 -o gpurun_out/r02a_regress -f \
    python tools/bench_membound.py > /dev/null 2>&1
ls -la gpurun_out | tail -12
