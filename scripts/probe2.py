"""Timing probe across tuning knobs (device-resident, CUDA events on the launch stream)."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from clawker_b200 import capture, engine, abi
n = int(os.environ.get("N", 10_000_000))
w = capture.make_workload(os.environ.get("CFG", "c2"))
pk = capture.gen_slots(w, min(n, 2_000_000), seed=2)
d = torch.from_numpy(pk).cuda()
d = d.repeat((n + pk.shape[0] - 1) // pk.shape[0], 1)[:n].contiguous()
v = torch.empty(n, dtype=torch.int32, device="cuda")
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream
for stages, minb, flags in itertools.product((1, 2), (1, 2), (0, 2)):
    if stages == 2 and minb == 2: continue
    os.environ["CLK_STAGES"] = str(stages); os.environ["CLK_MINB"] = str(minb)
    e = engine.Engine(device=0, max_metrics=1 << 20)
    e.load_tables(w.tables); e.compile_rules(w.rules); e.publish()
    for _ in range(3): e.classify_slots_dev(d.data_ptr(), 64, n, flags, v.data_ptr(), st)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5): e.classify_slots_dev(d.data_ptr(), 64, n, flags, v.data_ptr(), st)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 5; s = e.stats()
    print(f"stages={stages} minb={minb} flags={flags} ms={ms:.3f} Gpps={n/ms/1e6:.2f} GB/s={n*68/ms/1e6:.0f} grid={s.last_grid} smem={s.last_smem_bytes}", flush=True)
    e.close()
