"""Aggregate an ncu SASS profile by source function. usage: sass_by_func.py <cubin> <kernel-substr> <ncu_source.csv> <n_tiles>"""
import re,csv,collections,subprocess,os,sys
cubin,kern,prof_csv=sys.argv[1:4]; ntiles=float(sys.argv[4]) if len(sys.argv)>4 else 1
root=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dis=subprocess.run(["nvdisasm","-g","-c",cubin],capture_output=True,text=True).stdout.split("\n")
starts = [i for i, l in enumerate(dis) if l.startswith("\t.section") and kern in l and ".text." in l]
n_prof = sum(1 for r in csv.reader(open(sys.argv[3])) if r and r[0].startswith("0x"))
def n_inst(st):
    en = next((i for i in range(st + 1, len(dis)) if dis[i].startswith("\t.section")), len(dis))
    return sum(1 for l in dis[st:en] if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l))
# several instantiations may match the substring: the profiled one has as many instructions as the profile has rows (the page may list it twice)
start = min(starts, key=lambda st: min(abs(n_inst(st) - n_prof), abs(2 * n_inst(st) - n_prof)))
end=next((i for i in range(start+1,len(dis)) if dis[i].startswith("\t.section")),len(dis))
cur=None; insts=[]
for l in dis[start:end]:
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m: cur=(os.path.basename(m.group(1)),int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/",l): insts.append(cur)
rows=list(csv.reader(open(prof_csv))); hdr=rows[1]; ie=hdr.index("Instructions Executed"); at=hdr.index("Avg. Threads Executed"); ss=hdr.index("# Samples")
prof=[]
for r in rows[2:]:
    try: prof.append((int(r[ie]),float(r[at]),int(r[ss])))
    except: pass
marks={}
for f in os.listdir(os.path.join(root,"clawker_b200","csrc")):
    src=open(os.path.join(root,"clawker_b200","csrc",f)).read().split('\n'); mk=[]
    for i,l in enumerate(src,1):
        if l.startswith('__device__') or l.startswith('__global__') or l.startswith('CLK_HD') or re.match(r'\s+__device__ __forceinline__',l):
            nm=re.search(r'(\w+)\(',l); mk.append((i,nm.group(1) if nm else l[:30]))
    marks[f]=mk
def func_of(f,line):
    name=f
    for ln,nm in marks.get(f,[]):
        if ln<=line: name=nm
        else: break
    return name
agg=collections.defaultdict(lambda:[0,0.0,0])
for c,(n,a,s) in zip(insts,prof):
    key=func_of(*c) if c else '?'
    agg[key][0]+=n; agg[key][1]+=n*a; agg[key][2]+=s
tot=sum(v[0] for v in agg.values()); ts=sum(v[2] for v in agg.values())
print(f"total warp-inst {tot}  per-tile {tot/ntiles:.1f}")
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:32]: print(f"{k:28s} {100*v[0]/tot:5.1f}% inst thr={v[1]/max(v[0],1):5.1f} samp={100*v[2]/max(ts,1):5.1f}% per-tile={v[0]/ntiles:6.1f}")
