#!/bin/bash
# One GPU session: full GPU test-suite, smoke, the default bench line, the ncu launch list of the same step and one
# `ncu --set full` capture of the hash / MLP / Adam kernels (everything lands in gpurun_out/; TAG names the round).
# Afterwards, here:  python scripts/ncu_traffic.py gpurun_out/${TAG}_full_raw.csv profiles/ncu_traffic.json profiles/${TAG}_ncu_summary.csv
set -u
TAG=${TAG:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_1gpu.json
timeout 600 python bench.py --camera-optimizer off --no-cpu-baseline --no-eval > gpurun_out/${TAG}_bench_1gpu_camopt_off.json 2>/dev/null; cut -c1-200 gpurun_out/${TAG}_bench_1gpu_camopt_off.json
# launch list: every kernel of two eager steps at optimisation step 8 (true durations; eager event timings inflate small kernels)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --engine eager --steps 2 --warmup 3 --windows 1 --state-step 8 --no-cpu-baseline --no-eval \
    > gpurun_out/${TAG}_bench_under_ncu_list.log 2>&1; echo "ncu list rc=$?"
if [ "${FULL_NCU:-1}" = "1" ]; then
  # 16 matching launches per step (2 mlp_tc_pack, 2+2 density_fused, hashgrid fwd/dx/bwd, 2+2 mlp_tc, 2 live_compact, adam)
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:'hashgrid|density_fused|mlp_tc|adam|live_compact' --launch-skip 128 -c 32 -f -o gpurun_out/${TAG}_full \
      python bench.py --engine eager --steps 2 --warmup 3 --windows 1 --state-step 8 --no-cpu-baseline --no-eval \
      > gpurun_out/${TAG}_bench_under_ncu_full.log 2>&1; echo "ncu full rc=$?"
  ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
  rm -f gpurun_out/${TAG}_full.ncu-rep   # ~80 MB: gpurun only merges gpurun_out/ back when it is below 64 MiB
fi
ls -la gpurun_out | tail -12
