"""Launch the slot classifier a few times in one configuration (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from clawker_b200 import capture, engine, abi
n = int(os.environ.get("N", 10_000_000)); flags = int(os.environ.get("FLAGS", 0)); reps = int(os.environ.get("REPS", 3))
cfgname = os.environ.get("CFG", "c2")
w = capture.make_workload(cfgname)
pk = capture.gen_slots(w, min(n, 2_000_000), seed=2)
d = torch.from_numpy(pk).cuda()
d = d.repeat((n + pk.shape[0] - 1) // pk.shape[0], 1)[:n].contiguous()
v = torch.empty(n, dtype=torch.int32, device="cuda")
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
e = engine.Engine(device=0, max_metrics=1 << 20)
e.load_tables(w.tables); e.compile_rules(w.rules); e.publish()
for _ in range(reps):
    e.classify_slots_dev(d.data_ptr(), 64, n, flags, v.data_ptr(), ts.cuda_stream)
torch.cuda.synchronize()
print("done", e.stats().last_grid)
