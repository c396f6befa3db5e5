"""DRAM bytes per batch from an ncu launch list (the --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum pass)
-> profiles/<tag>_varlen_traffic.json (the varlen pipeline, one whole batch = the first run of consecutive slices) and
profiles/<tag>_traffic.json (k_slots_fast, one launch).  bench.py reads both for roofline.traffic.
usage: traffic_from_launches.py <tag> <launches.csv> <packets per varlen batch> <packets per slots launch>"""
import collections, csv, json, os, sys
tag, path, n_var, n_slot = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
h = rows[0]; ID, K, M, U, V = (h.index(x) for x in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
launch = collections.OrderedDict()
for r in rows[1:]:
    d = launch.setdefault(int(r[ID]), {"k": r[K].split("(")[0].strip()})
    v = float(r[V].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[U], 1)
    d[r[M]] = v
seq = list(launch.values())
is_var = lambda k: "k_varlen_head" in k or "k_name_list" in k or "k_generic_list" in k
# one batch of the varlen path: from the first k_varlen_head to the launch before the (n_slices + 1)-th head
heads = [i for i, d in enumerate(seq) if "k_varlen_head" in d["k"]]
slice_pk = 12582912
n_slices = -(-n_var // slice_pk)
first = heads[0]; last = heads[n_slices] if len(heads) > n_slices else len(seq)
per = collections.OrderedDict(); tot = 0.0
for d in seq[first:last]:
    if not is_var(d["k"]): continue
    b = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    a = per.setdefault(d["k"], {"launches": 0, "us": 0.0, "dram_bytes": 0.0}); a["launches"] += 1; a["us"] += d.get("gpu__time_duration.sum", 0); a["dram_bytes"] += b; tot += b
src = f"{os.path.basename(path)} (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none)"
json.dump({"kernel": f"varlen pipeline (k_varlen_head + k_generic_list + k_name_list<1,3,4>), {n_slices} slices of {slice_pk} packets", "packets_per_launch": n_var,
           "dram_bytes_per_launch": tot, "dram_bytes_per_packet": round(tot / n_var, 2), "per_kernel": per, "algorithmic_bytes_per_packet": 124.24, "source": src},
          open(os.path.join(root, "profiles", f"{tag}_varlen_traffic.json"), "w"), indent=1)
print("varlen: %.1f B per packet, %.0f us under ncu" % (tot / n_var, sum(a["us"] for a in per.values())))
sl = [d for d in seq if "k_slots_fast" in d["k"]]
if sl:
    mx = max(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in sl)
    full = [d for d in sl if d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] >= 0.9 * mx]        # the whole-batch launches (e2e / sustained legs launch chunks)
    rd = sum(d["dram__bytes_read.sum"] for d in full) / len(full); wr = sum(d["dram__bytes_write.sum"] for d in full) / len(full)
    json.dump({"kernel": "clk::k_slots_fast", "packets_per_launch": n_slot, "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
               "algorithmic_bytes_per_launch": n_slot * 68.0, "ncu_duration_us": sum(d["gpu__time_duration.sum"] for d in full) / len(full),
               "launches_averaged": len(full), "source": src},
              open(os.path.join(root, "profiles", f"{tag}_traffic.json"), "w"), indent=1)
    print("slots: %.1f B per packet over %d launches" % ((rd + wr) / n_slot, len(full)))
