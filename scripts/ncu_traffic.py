"""Summarise an `ncu --set full` raw-page CSV into profiles/ncu_dram_traffic.json (DRAM bytes per launch of the hash /
MLP / Adam kernels), keyed the way bench.py names its C-ABI calls, plus a small per-kernel table for profiles/.

    ncu -i gpurun_out/r01_final.ncu-rep --page raw --csv > gpurun_out/r01_final_raw.csv
    python scripts/ncu_traffic.py gpurun_out/r01_final_raw.csv profiles/ncu_dram_traffic.json profiles/r01_ncu_final_summary.csv
"""
import csv, json, sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3,
        "nsecond": 1e-3}
RAYS = 4096
# kernel-name fragment -> (C-ABI call, candidate point counts, largest first)
MAP = [("hashgrid_fwd_kernel", "b2n_hashgrid_fwd", [RAYS * 48]), ("hashgrid_bwd", "b2n_hashgrid_bwd", [RAYS * 48]),
       ("density_fused_fwd", "b2n_density_field_fwd", [RAYS * 256, RAYS * 96]),
       ("density_fused_bwd", "b2n_density_field_bwd", [RAYS * 256, RAYS * 96]),
       ("mlp_tc_fwd", "b2n_mlp_tc_fwd", None), ("mlp_tc_bwd", "b2n_mlp_tc_bwd", None), ("adam", "b2n_adam_step_dev", None)]
COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def main(raw, out_json, out_csv):
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {k: i for i, k in enumerate(hdr)}

    def val(r, k):
        if k not in ix or r[ix[k]] in ("", "n/a"):
            return None
        return float(r[ix[k]].replace(",", "")) * UNIT.get(units[ix[k]], 1.0)

    groups = {}
    for r in data:
        name = r[ix["Kernel Name"]]
        for frag, abi, _ in MAP:
            if frag in name:
                groups.setdefault(abi, []).append(r)
                break
    traffic, table = {}, []
    for frag, abi, ns in MAP:
        rs = groups.get(abi, [])
        if not rs:
            continue
        # the same kernel at different sizes: the longer launch is the larger point count
        uniq = {}
        for r in rs:
            dyn = r[ix["launch__shared_mem_per_block_dynamic"]] if "launch__shared_mem_per_block_dynamic" in ix else ""
            uniq.setdefault((r[ix["Grid Size"]], r[ix["Block Size"]], dyn), r)
        ordered = sorted(uniq.values(), key=lambda r: -(val(r, "gpu__time_duration.sum") or 0))
        for i, r in enumerate(ordered):
            key = f"{abi}[n={ns[i]}]" if ns and i < len(ns) else f"{abi}#{i}"
            rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
            traffic.setdefault(abi, {})[key] = None if rd is None else rd + wr
            table.append([key, r[ix["Kernel Name"]][:50], r[ix["Grid Size"]], r[ix["Block Size"]]] + [val(r, c) for c in COLS])
    json.dump(traffic, open(out_json, "w"), indent=1)
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["call", "kernel", "grid", "block"] + [c + (" [us]" if "time" in c else " [B]" if "bytes" in c else "") for c in COLS])
        w.writerows(table)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:4])
