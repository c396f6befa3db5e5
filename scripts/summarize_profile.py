"""Turn gpurun_out/<name>.ncu-rep (+ launch list csv) into committed evidence under profiles/.
usage: summarize_profile.py <tag e.g. r01 or r01_varlen> <ncu-rep> <kernel substring> <packets per launch> [launch csv] [algorithmic bytes per packet]"""
import csv, json, subprocess, sys, os, collections, io
tag, rep, kern, npk = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
launch_csv = sys.argv[5] if len(sys.argv) > 5 and sys.argv[5] != '-' else None
bpp = float(sys.argv[6]) if len(sys.argv) > 6 else 68.0
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(root, "profiles"); os.makedirs(out_dir, exist_ok=True)
kfilter = ["--kernel-name", "regex:" + kern.split("::")[-1].split("<")[0].split(" ")[0], "--launch-count", "1"]      # a report may hold several kernels: the first launch of this one
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + kfilter, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, vals = rows[0], rows[1], rows[2]
def g(name):
    return (vals[hdr.index(name)], units[hdr.index(name)]) if name in hdr else (None, None)
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
rd, wr = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
dur, du = g("gpu__time_duration.sum")
dur_us = float(dur.replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3}.get(du, 1)
traffic = to_bytes(*rd) + to_bytes(*wr)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + kfilter, capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src))); sh = srows[1]
cols = [i for i, h in enumerate(sh) if h.startswith("stall_") and "Not Issued" not in h]
st = collections.Counter()
ie, at = sh.index("Instructions Executed"), sh.index("Avg. Threads Executed"); n_inst = 0
body = srows[2:]
for j, r in enumerate(body):                  # (the source page repeats the listing once per view: keep the first copy)
    if j and r[:1] == body[0][:1] and r[1:2] == body[0][1:2]: body = body[:j]; break
for r in body:
    for i in cols:
        try: st[sh[i]] += int(r[i])
        except Exception: pass
    try: n_inst += int(r[ie])
    except Exception: pass
tot = sum(st.values()) or 1
lines = [f"# {tag}: ncu --set full --clock-control none, kernel `{kern}`, {npk} packets per launch", "",
         "Per-launch values are cold-cache and serialised by ncu's replay: use shares and traffic, not the absolute time.", "",
         "| metric | value | unit |", "|---|---|---|"]
for w in want:
    v, u = g(w)
    if v is not None: lines.append(f"| {w} | {v} | {u} |")
lines += ["", f"- DRAM traffic per launch: {traffic/1e6:.1f} MB (read {to_bytes(*rd)/1e6:.1f} + write {to_bytes(*wr)/1e6:.1f}); "
          f"algorithmic bytes {npk*bpp/1e6:.1f} MB ({bpp:g} B/packet) -> traffic/algorithmic = {traffic/(npk*bpp):.3f}",
          f"- duration under ncu {dur_us:.1f} us -> {npk*bpp/dur_us/1e3:.0f} GB/s algorithmic (cold cache; the bench's CUDA-event number is the one to quote)",
          f"- warp-instructions per 32-packet tile: {n_inst/(npk/32):.0f}", "", "Warp stall samples:", ""]
for k, v in st.most_common(10): lines.append(f"- {k}: {100*v/tot:.1f}%")
if launch_csv:
    lr = [r for r in csv.reader(open(launch_csv)) if len(r) > 5]
    lh = lr[0]; ki, vi, ui = lh.index("Kernel Name"), lh.index("Metric Value"), lh.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in lr[1:]:
        try: v = float(r[vi].replace(",", ""))
        except Exception: continue
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(r[ui], 1)
        a = agg.setdefault(r[ki].split("(")[0][:80], [0, 0.0]); a[0] += 1; a[1] += v
    t = sum(a[1] for a in agg.values())
    lines += ["", f"Launch list of `bench.py --steps 5 --warmup 3` (`{os.path.basename(launch_csv)}`, gpu__time_duration.sum):", "",
              "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, a in agg.items(): lines.append(f"| {k} | {a[0]} | {a[1]:.1f} | {100*a[1]/t:.1f}% |")
    import shutil; shutil.copy(launch_csv, os.path.join(out_dir, f"{tag}_launches.csv"))
open(os.path.join(out_dir, f"{tag}_summary.md"), "w").write("\n".join(lines) + "\n")
json.dump({"kernel": kern, "packets_per_launch": npk, "dram_bytes_per_launch": traffic, "dram_read": to_bytes(*rd), "dram_write": to_bytes(*wr),
           "algorithmic_bytes_per_launch": npk * bpp, "ncu_duration_us": dur_us, "source": os.path.basename(rep)},
          open(os.path.join(out_dir, f"{tag}_traffic.json"), "w"), indent=1)
print("\n".join(lines))
