"""Condense an .ncu-rep into the few numbers DESIGN.md / bench.py cite (run here, no GPU needed).
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/out_prefix"""
import csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
summary = []
for r in rows[2:]:
    k = {"kernel": r[idx["Kernel Name"]]}
    for m in KEEP:
        if m in idx:
            k[m] = f"{r[idx[m]]} {units[idx[m]]}".strip()
    stalls = []
    for h in hdr:
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            try:
                stalls.append((int(float(r[idx[h]])), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(s for s, _ in stalls) or 1
    k["top_stalls_pct"] = {n: round(100.0 * s / tot, 1) for s, n in sorted(stalls, reverse=True)[:6]}
    def num(m):
        v = r[idx[m]].replace(",", "")
        u = units[idx[m]]
        f = float(v)
        return f * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(u, 1)
    k["dram_traffic_bytes"] = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    summary.append(k)
json.dump(summary, open(out + ".json", "w"), indent=1)
with open(out + ".txt", "w") as f:
    for k in summary:
        f.write(k["kernel"] + "\n")
        for m, v in k.items():
            if m != "kernel":
                f.write(f"    {m:75s} {v}\n")
print(open(out + ".txt").read()[:3000])
