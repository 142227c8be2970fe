"""Pick the judged metrics out of an ncu report: python scripts/ncu_extract.py X.ncu-rep [--every 3] > profiles/X.json"""
import csv, json, subprocess, sys
rep = sys.argv[1]
every = int(sys.argv[sys.argv.index("--every") + 1]) if "--every" in sys.argv else 1
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keep = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
stall = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
out = []
for i, r in enumerate(data):
    if (i + 1) % every:
        continue
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    e = {k: (d[k] + (" " + u[k] if u.get(k) else "")) for k in keep if k in d and d[k] != ""}
    s = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): int(float(d[h])) for h in stall if d.get(h)}
    tot = sum(s.values()) or 1
    e["warp_stall_samples_pct"] = {k: round(100.0 * v / tot, 1) for k, v in sorted(s.items(), key=lambda kv: -kv[1])[:8]}
    out.append(e)
print(json.dumps(out, indent=1))
