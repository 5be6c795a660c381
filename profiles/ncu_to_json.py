"""Condense `ncu -i X.ncu-rep --page raw --csv` exports into profiles/r1_ncu_kernels.json (per kernel: duration,
DRAM bytes, pipe utilisation, registers, stall mix).  usage: python profiles/ncu_to_json.py out.json raw1.csv [raw2.csv ...]"""
import csv, json, re, sys

KEEP = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_active_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed": "dram_busy_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefronts_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
}
UNIT = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}
out = {}
for path in sys.argv[2:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("<unnamed>::", "").replace("void ", "")
        d = {}
        for i, h in enumerate(hdr):
            if h in KEEP and r[i]:
                v = float(r[i].replace(",", ""))
                d[KEEP[h]] = v * UNIT.get(units[i], 1.0) if KEEP[h] in ("duration_us", "dram_read", "dram_write") else v
            if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and r[i] and float(r[i]) >= 0.2:
                d.setdefault("stalls_per_issue", {})[h.split("stalled_")[1].split("_per_issue")[0]] = round(float(r[i]), 2)
        d["dram_bytes"] = d.get("dram_read", 0) + d.get("dram_write", 0)
        tag = path.split("/")[-1].replace("_raw.csv", "").replace("full_", "")
        out[f"{name} [{tag}]"] = d
json.dump(out, open(sys.argv[1], "w"), indent=1, sort_keys=True)
print(json.dumps({k: (v["duration_us"], v["dram_bytes"]) for k, v in out.items()}))
