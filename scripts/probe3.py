"""Varlen + name-stage timing probe (BASELINE config 3 shape), device-resident."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from clawker_b200 import capture, engine, abi
gen = int(os.environ.get("GEN", 1_000_000)); reps = int(os.environ.get("REPS", 20))
w = capture.make_workload("c3")
blob, offs, info = capture.gen_varlen(w, gen, seed=3, pool=8192)
print(info, "avg B/pkt", blob.size / gen, flush=True)
db = torch.from_numpy(blob).cuda().repeat(reps)
o64 = torch.from_numpy(offs.astype(np.int64)).cuda()
per = int(offs[-1])
allo = torch.cat([o64[:-1] + k * per for k in range(reps)] + [o64[-1:] + (reps - 1) * per]).to(torch.int32)   # fits u32 (16 B units)
n = gen * reps
v = torch.empty(n, dtype=torch.int32, device="cuda")
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
e = engine.Engine(device=0, max_metrics=1 << 20, max_dns_cache=1 << 21, max_routes=1 << 16)
e.load_tables(w.tables); e.compile_rules(w.rules); e.publish()
for flags in (abi.F_NAME_STAGE, 0, abi.F_NAME_STAGE | abi.F_NO_METRICS):
    for _ in range(2): e.classify_varlen_dev(db.data_ptr(), allo.data_ptr(), n, flags, v.data_ptr(), ts.cuda_stream)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(3): e.classify_varlen_dev(db.data_ptr(), allo.data_ptr(), n, flags, v.data_ptr(), ts.cuda_stream)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 3
    byts = db.numel() + 8 * n
    print(f"flags={flags} n={n} ms={ms:.3f} Gpps={n/ms/1e6:.2f} GB/s={byts/ms/1e6:.0f} grid={e.stats().last_grid}", flush=True)
