"""Join an `ncu --page source --csv` dump with nvdisasm line info -> source lines ranked by warp-stall samples.
usage: stall_by_line.py <cubin> <kernel-substring> <ncu_source.csv> [top]"""
import re, csv, collections, subprocess, os, sys
cubin, kern, prof = sys.argv[1:4]; top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
starts = [i for i, l in enumerate(dis) if l.startswith("\t.section") and kern in l and ".text." in l]
n_prof = sum(1 for r in csv.reader(open(sys.argv[3])) if r and r[0].startswith("0x"))
def n_inst(st):
    en = next((i for i in range(st + 1, len(dis)) if dis[i].startswith("\t.section")), len(dis))
    return sum(1 for l in dis[st:en] if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l))
# several instantiations may match the substring: the profiled one has as many instructions as the profile has rows (the page may list it twice)
start = min(starts, key=lambda st: min(abs(n_inst(st) - n_prof), abs(2 * n_inst(st) - n_prof)))
end = next((i for i in range(start + 1, len(dis)) if dis[i].startswith("\t.section")), len(dis))
cur = None; insts = []
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): insts.append((cur, l.strip()[:90]))
rows = list(csv.reader(open(prof))); hdr = rows[1]
ie, ss, at = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: [0, 0, collections.Counter(), 0.0]); tot = 0; toti = 0
for (c, txt), r in zip(insts, rows[2:]):
    try: n = int(r[ie]); s = int(r[ss]); a = float(r[at])
    except Exception: continue
    k = c or ("?", 0); agg[k][0] += n; agg[k][1] += s; agg[k][3] += n * a; tot += s; toti += n
    for i in stall_cols:
        try: agg[k][2][hdr[i]] += int(r[i])
        except Exception: pass
print(f"sass {len(insts)} rows {len(rows)-2} total warp-inst {toti} samples {tot}")
src = {}; d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "clawker_b200", "csrc")
for f in os.listdir(d):
    if os.path.isfile(os.path.join(d, f)): src[f] = open(os.path.join(d, f)).read().split("\n")
allst = collections.Counter()
for v in agg.values(): allst.update(v[2])
print("stalls:", ", ".join(f"{a.replace('stall_','')} {100*b/max(1,sum(allst.values())):.1f}%" for a, b in allst.most_common(8)))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    t = src[k[0]][k[1] - 1].strip()[:64] if k[0] in src and 0 < k[1] <= len(src[k[0]]) else ""
    st = ", ".join(f"{a.replace('stall_','')}:{100*b/max(1,v[1]):.0f}%" for a, b in v[2].most_common(2))
    print(f"{100*v[1]/tot:5.1f}% samp inst={100*v[0]/toti:4.1f}% thr={v[3]/max(1,v[0]):4.1f} {k[0]}:{k[1]} {t} | {st}")
