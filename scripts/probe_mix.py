"""Slot kernel time vs share of rare frames (IPv6 / non-IP)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from clawker_b200 import capture, engine
n = 10_000_000
w = capture.make_workload("c2")
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream
e = engine.Engine(device=0, max_metrics=131072); e.load_tables(w.tables); e.compile_rules(w.rules); e.publish()
for v6, nonip in ((0.0, 0.0), (0.0, 0.005), (0.005, 0.0), (0.02, 0.005), (0.05, 0.0)):
    pk = capture.gen_slots(w, 2_000_000, seed=2, v6_frac=v6, nonip_frac=nonip)
    d = torch.from_numpy(pk).cuda().repeat(5, 1).contiguous(); v = torch.empty(n, dtype=torch.int32, device="cuda")
    for flags in (0, 2):
        for _ in range(3): e.classify_slots_dev(d.data_ptr(), 64, n, flags, v.data_ptr(), st)
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(5): e.classify_slots_dev(d.data_ptr(), 64, n, flags, v.data_ptr(), st)
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 5
        print(f"v6={v6} nonip={nonip} flags={flags} ms={ms:.3f} Gpps={n/ms/1e6:.2f}", flush=True)
