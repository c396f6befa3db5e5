"""Throughput vs batch size (fixed per-launch cost vs steady state)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from clawker_b200 import capture, engine, abi
w = capture.make_workload("c2")
pk = capture.gen_slots(w, 2_000_000, seed=2)
base = torch.from_numpy(pk).cuda()
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream
e = engine.Engine(device=0, max_metrics=131072)
e.load_tables(w.tables); e.compile_rules(w.rules); e.publish()
for n in (1_000_000, 2_000_000, 5_000_000, 10_000_000, 20_000_000, 50_000_000, 100_000_000):
    d = base.repeat((n + 1_999_999) // 2_000_000, 1)[:n].contiguous(); v = torch.empty(n, dtype=torch.int32, device="cuda")
    for flags in (0, 2):
        for _ in range(3): e.classify_slots_dev(d.data_ptr(), 64, n, flags, v.data_ptr(), st)
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(); K = 5
        for _ in range(K): e.classify_slots_dev(d.data_ptr(), 64, n, flags, v.data_ptr(), st)
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / K
        print(f"n={n:>10} flags={flags} ms={ms:8.3f} Gpps={n/ms/1e6:6.2f} frac={n*68/ms/1e6/6572.2:.3f}", flush=True)
    del d, v
