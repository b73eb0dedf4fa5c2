"""Condense `ncu --page raw --csv` into a small JSON/markdown summary (one entry per kernel: median of launches)."""
import csv
import json
import statistics
import sys

KEYS = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_throughput_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "sm__maximum_warps_per_active_cycle_pct": "theoretical_occupancy_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "lts__t_sectors_srcunit_tex_op_read.sum": "l2_read_sectors",
    "lts__t_sectors_srcunit_tex_op_write.sum": "l2_write_sectors",
    "lts__t_sectors_srcunit_tex_op_red.sum": "l2_red_sectors",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_bytes",
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "Kbyte/block": 1e3, "byte/block": 1.0}


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").strip()
        d = per.setdefault(name, {})
        for k, short in KEYS.items():
            if k in idx and r[idx[k]] not in ("", "n/a"):
                try:
                    v = float(r[idx[k]].replace(",", "")) * UNIT_SCALE.get(units[idx[k]], 1.0)
                except ValueError:      # "no data"
                    continue
                d.setdefault(short, []).append(v)
    out = {}
    for name, d in per.items():
        e = {k: statistics.median(v) for k, v in d.items()}
        e["launches"] = len(next(iter(d.values())))
        if "dram_read" in e and "dram_write" in e:
            e["dram_traffic_bytes"] = e["dram_read"] + e["dram_write"]
        out[name] = e
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
