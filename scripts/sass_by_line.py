"""Join an `ncu --page source --csv` SASS dump with nvdisasm line info -> hot source lines.
usage: sass_by_line.py <cubin> <kernel-substring> <ncu_source.csv> [top]"""
import re, csv, collections, sys, subprocess, os
cubin, kern, prof_csv = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
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
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        insts.append((cur, l.strip()[:80]))
rows = list(csv.reader(open(prof_csv)))
hdr = rows[1]; ie = hdr.index("Instructions Executed"); at = hdr.index("Avg. Threads Executed"); ss = hdr.index("# Samples")
prof = []
for r in rows[2:]:
    try: prof.append((int(r[ie]), float(r[at]), int(r[ss])))
    except Exception: pass
print("sass insts", len(insts), "profile rows", len(prof))
agg = collections.defaultdict(lambda: [0, 0, 0])
for (c, _), (n, a, s) in zip(insts, prof):
    k = c or ("?", 0); agg[k][0] += n; agg[k][1] += n * a; agg[k][2] += s
tot = sum(v[0] for v in agg.values()); tots = sum(v[2] for v in agg.values())
src = {}
d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "clawker_b200", "csrc")
for f in [x for x in os.listdir(d) if os.path.isfile(os.path.join(d, x))]:
    src[f] = open(os.path.join(d, f)).read().split("\n")
print(f"total warp-inst {tot}, samples {tots}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    t = src[k[0]][k[1] - 1].strip()[:90] if k[0] in src and 0 < k[1] <= len(src[k[0]]) else ""
    print(f"{100*v[0]/tot:5.1f}% thr={v[1]/max(v[0],1):5.1f} samp={100*v[2]/max(tots,1):4.1f}% {k[0]}:{k[1]} {t}")
