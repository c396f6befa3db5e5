"""Every kernel of an .ncu-rep (ncu --set full) -> one markdown file under profiles/ with a metric table per kernel.
usage: summarize_kernels.py <out name> <ncu-rep> <title> [kernel-substring=packets ...]   (packets -> warp-instructions per packet)"""
import csv, io, os, subprocess, sys
out, rep, title = sys.argv[1:4]
pk = dict(a.rsplit("=", 1) for a in sys.argv[4:] if "=" in a)
only = [a for a in sys.argv[4:] if "=" not in a]                       # plain arguments: keep only kernels whose name contains one of them
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum"]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") or h.startswith("smsp__average_warp_latency_issue_stalled_")]
lines = [f"# {title}", "", "Per-launch values are cold-cache and serialised by ncu's replay: use shares and ratios, not the absolute time.", ""]
seen = set()
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("clk::", "")
    if name in seen or (only and not any(o in name for o in only)): continue
    seen.add(name)
    lines += [f"## {name}", "", "| metric | value | unit |", "|---|---|---|"]
    for w in want:
        if w in hdr: lines.append(f"| {w} | {r[hdr.index(w)]} | {units[hdr.index(w)]} |")
    for k, n in pk.items():
        if k in name: lines.append(f"| warp-instructions per packet ({n} packets) | {float(r[hdr.index('smsp__inst_executed.sum')].replace(',', '')) / float(n):.1f} | |")
    st = []
    for h in stalls:
        try: st.append((float(r[hdr.index(h)].replace(",", "")), h.split("issue_stalled_")[1].split("_per_")[0].replace(".ratio", "")))
        except Exception: pass
    st.sort(reverse=True)
    if st: lines += ["", "Warp stalls per issue-active cycle: " + ", ".join(f"{n} {v:.2f}" for v, n in st[:6])]
    lines.append("")
open(os.path.join(root, "profiles", out), "w").write("\n".join(lines))
print("\n".join(lines))
