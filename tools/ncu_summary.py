"""Summarise an .ncu-rep (ncu --set full) into JSON: per kernel duration, DRAM bytes, issue /
DRAM / occupancy percentages.  usage: ncu_summary.py report.ncu-rep out.json"""
import csv, io, json, re, subprocess, sys
rep, out = sys.argv[1:3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "launch__registers_per_thread": "registers",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_smem_ctas",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "launch__grid_size": "grid", "launch__block_size": "block",
}
units = rows[1]
res = {}
for r in rows[2:]:
    if len(r) < len(h):
        continue
    name = re.sub(r"\(.*", "", r[h.index("Kernel Name")]).replace("void ", "").replace("chips::", "")
    d = {}
    for i, col in enumerate(h):
        if col in want:
            v = float(r[i].replace(",", "")) if r[i] else None
            u = units[i]
            if col == "gpu__time_duration.sum":
                v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
            if col.startswith("dram__bytes"):
                v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            d[want[col]] = v
    d["dram_bytes"] = (d.get("dram_read_bytes") or 0) + (d.get("dram_write_bytes") or 0)
    res.setdefault(name, d)
json.dump(res, open(out, "w"), indent=1)
for k, v in res.items():
    print(f"{k:40s} {v.get('duration_us', 0):9.1f} us  dram {v['dram_bytes']/1e6:8.1f} MB  issue {v.get('issue_active_pct')}%  warps {v.get('warps_active_pct')}%")
