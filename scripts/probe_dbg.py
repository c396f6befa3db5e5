import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from clawker_b200 import capture, engine
n = 20_000_000
w = capture.make_workload("c2"); pk = capture.gen_slots(w, 2_000_000, seed=2)
d = torch.from_numpy(pk).cuda().repeat(10, 1).contiguous(); v = torch.empty(n, dtype=torch.int32, device="cuda")
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream
e = engine.Engine(device=0, max_metrics=131072); e.load_tables(w.tables); e.compile_rules(w.rules); e.publish()
for name, flags in (("full", 0), ("match only", 0x100), ("global only", 0x200), ("cta only", 0x300), ("no metrics", 2)):
    for _ in range(3): e.classify_slots_dev(d.data_ptr(), 64, n, flags, v.data_ptr(), st)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5): e.classify_slots_dev(d.data_ptr(), 64, n, flags, v.data_ptr(), st)
    t1.record(); torch.cuda.synchronize()
    print(f"{name:12s} ms={t0.elapsed_time(t1)/5:.3f}", flush=True)
