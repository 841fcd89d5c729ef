"""Summarise an `ncu --set full` raw-page CSV of ONE eager nerfacto step into profiles/ncu_traffic.json — per-launch DRAM
bytes (dram__bytes_read.sum + dram__bytes_write.sum), L2 bytes (lts__t_sectors.sum x 32 B), duration and tensor-pipe
activity of the hash / MLP / Adam kernels, keyed exactly as bench.py names its C-ABI calls — plus a per-kernel table.

    ncu --set full --clock-control none --import-source on -k regex:'hashgrid|density_fused|mlp_tc|adam|live_compact' \
        --launch-skip 128 -c 32 -f -o gpurun_out/r02_full \
        python bench.py --engine eager --steps 2 --warmup 3 --windows 1 --state-step 8 --no-cpu-baseline --no-eval   (scripts/gpu_final.sh)
    ncu -i gpurun_out/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv
    python scripts/ncu_traffic.py gpurun_out/r02_full_raw.csv profiles/ncu_traffic.json profiles/r02_ncu_summary.csv

Launches of one kernel are told apart by their order inside the step (engine.NerfactoStep._body issues a fixed sequence):
density_fused_fwd level 0 (n = 4096 x 256) then level 1 (n = 4096 x 96); mlp_tc_fwd base (32 -> 16) then head (63 -> 3);
mlp_tc_bwd head then base; density_fused_bwd level 0 then level 1.  Only the LAST captured step is used.
"""
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3,
        "nsecond": 1e-3, "sector": 1.0}
R = 4096
ORDER = [  # (kernel-name fragment, [bench keys in launch order within a step])
    ("density_fused_fwd", [f"b2n_density_field_fwd[n={R * 256}]", f"b2n_density_field_fwd[n={R * 96}]"]),
    ("hashgrid_fwd", [f"b2n_hashgrid_fwd[n={R * 48}]"]),
    ("mlp_tc_fwd", [f"b2n_mlp_tc_fwd[n={R * 48},in=32,out=16]", f"b2n_mlp_tc_fwd[n={R * 48},in=63,out=3]"]),
    ("mlp_tc_bwd", [f"b2n_mlp_tc_bwd[n={R * 48},in=63,out=3]", f"b2n_mlp_tc_bwd[n={R * 48},in=32,out=16]"]),
    ("hashgrid_bwd", [f"b2n_hashgrid_bwd[n={R * 48}]"]),
    ("hashgrid_dx", [f"b2n_hashgrid_dx[n={R * 48}]"]),
    ("density_fused_bwd", [f"b2n_density_field_bwd[n={R * 256}]", f"b2n_density_field_bwd[n={R * 96}]"]),
    ("adam", ["b2n_adam_step_dev"]),
]
COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]


def main(raw, out_json, out_csv):
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {k: i for i, k in enumerate(hdr)}

    def val(r, k):
        if k not in ix or r[ix[k]] in ("", "n/a"):
            return None
        try:
            return float(r[ix[k]].replace(",", "")) * UNIT.get(units[ix[k]], 1.0)
        except ValueError:
            return None

    out, table = {}, []
    for frag, keys in ORDER:
        rs = [r for r in data if frag in r[ix["Kernel Name"]]]
        if not rs:
            continue
        rs = rs[-len(keys):]  # the last captured step
        for key, r in zip(keys, rs):
            rd, wr, sec = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum"), val(r, "lts__t_sectors.sum")
            out[key] = {"kernel": r[ix["Kernel Name"]][:60], "time_us": val(r, "gpu__time_duration.sum"),
                        "dram_bytes": None if rd is None else rd + wr, "lts_bytes": None if sec is None else 32.0 * sec,
                        "tensor_pipe_active_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                        "l1tex_pct": val(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
                        "lts_pct": val(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                        "registers": val(r, "launch__registers_per_thread")}
            table.append([key, r[ix["Kernel Name"]][:60], r[ix["Grid Size"]], r[ix["Block Size"]]] + [val(r, c) for c in COLS])
    json.dump(out, open(out_json, "w"), indent=1)
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["call", "kernel", "grid", "block"] + COLS)
        w.writerows(table)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:4])
