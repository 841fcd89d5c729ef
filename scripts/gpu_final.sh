#!/bin/bash
# One GPU session: full GPU test-suite, smoke, the default bench line and the ncu launch list of the same command
# (everything lands in gpurun_out/).  `FULL_NCU=1` adds one `ncu --set full` capture of the hash / MLP / Adam kernels.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_final.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --engine eager --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_list.log 2>&1; echo "ncu list rc=$?"
if [ "${FULL_NCU:-0}" = "1" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'hashgrid|density_fused|mlp_tc|adam' -c 11 \
      -f -o gpurun_out/r01_final python bench.py --engine eager --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_full.log 2>&1; echo "ncu full rc=$?"
  ncu -i gpurun_out/r01_final.ncu-rep --page raw --csv > gpurun_out/r01_final_raw.csv 2>/dev/null
fi
ls -la gpurun_out | tail -8
