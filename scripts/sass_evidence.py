"""cuobjdump -sass of the shipped library -> profiles/<tag>_sass_evidence.md: a mnemonic census per kernel and the lines around every
TMA / mbarrier / bulk-copy / election instruction of the slot kernel.  CPU only.   usage: sass_evidence.py <tag> [lib.so]"""
import collections, os, re, subprocess, sys
tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(root, "clawker_b200", "libclawker_gpu.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("\n")
funcs = collections.OrderedDict(); cur = None
for l in sass:
    m = re.search(r"Function : (\S+)", l)
    if m: cur = m.group(1); funcs[cur] = []; continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): funcs[cur].append(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l.rstrip()))
names = subprocess.run(["cu++filt"] + list(funcs), capture_output=True, text=True).stdout.split("\n")
def short(n):
    m = re.match(r"(?:void )?([\w:]+(?:<[^>]*>)?)", n)
    return (m.group(1) if m else n).replace("(int)", "").replace("(bool)1", "true").replace("(bool)0", "false")
dem = {k: short(n) for k, n in zip(funcs, names)}
cols = ["UTMALDG", "UBLKCP", "SYNCS", "ELECT", "LDG.E.128", "LDG.E.CONSTANT", "LDS.128", "ATOMS", "ATOMG", "RED.E", "MATCH", "VOTE", "SHFL", "WARPSYNC", "R2UR"]
tensor = sum(1 for b in funcs.values() for l in b if re.search(r"\b(HMMA|IMMA|UTCMMA|UTCHMMA|QGMMA|HGMMA|TCGEN05)", l))
want = ["k_slots_fast<1, true, 1, true>", "k_varlen_head<16, 2>", "k_name_list<1, 3>", "k_name_list<3, 3>", "k_name_list<4, 3>", "k_name_list<5, 3>", "k_generic_list", "k_hdr16", "k_events",
        "k_generic<true>", "k_counters_totals", "k_counters_export"]
out = [f"# SASS evidence ({tag}) — `cuobjdump -sass` of the shipped clawker_b200/libclawker_gpu.so (sm_100a only; `scripts/sass_evidence.py`)", "",
       "Mnemonic census per kernel (whole function, callees included) and the lines around every TMA / mbarrier / bulk-copy / election instruction of the",
       f"slot kernel.  Tensor-core instructions (HMMA / IMMA / UTCMMA / tcgen05) in the whole library: {tensor} — this path has no dense contraction.", "",
       "| kernel | SASS instructions | " + " | ".join(cols) + " |", "|---|---|" + "---|" * len(cols)]
for w in want:
    for k, d in dem.items():
        if d.replace("clk::", "") == w:
            body = funcs[k]
            out.append(f"| `{d}` | {len(body)} | " + " | ".join(str(sum(1 for l in body if re.search(r"\b" + re.escape(c) + r"\b", l))) for c in cols) + " |")
key = next(k for k, d in dem.items() if d.replace("clk::", "") == want[0])
body = funcs[key]
hot = [i for i, l in enumerate(body) if re.search(r"UTMALDG|UBLKCP|SYNCS|ELECT|FENCE", l)]
keep = sorted({j for i in hot for j in (i - 1, i, i + 1) if 0 <= j < len(body)})
out += ["", f"## `{dem[key]}`: every UTMALDG / UBLKCP / SYNCS / ELECT / FENCE line (with the instruction before and after)", "```"]
prev = None
for j in keep:
    if prev is not None and j != prev + 1: out.append("        ...")
    out.append(body[j]); prev = j
out.append("```")
open(os.path.join(root, "profiles", f"{tag}_sass_evidence.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:24]))
