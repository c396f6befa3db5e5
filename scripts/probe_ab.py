"""A/B a list of library builds on the c2 slot workload: CLK_LIB=<so> per subprocess. usage: probe_ab.py a.so b.so ..."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch
from clawker_b200 import capture, engine
n = 10_000_000
w = capture.make_workload("c2")
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream
e = engine.Engine(device=0, max_metrics=131072); e.load_tables(w.tables); e.compile_rules(w.rules); e.publish()
for v6, nonip in ((0.0, 0.0), (0.02, 0.005)):
    pk = capture.gen_slots(w, 2_000_000, seed=2, v6_frac=v6, nonip_frac=nonip)
    d = torch.from_numpy(pk).cuda().repeat(5, 1).contiguous(); v = torch.empty(n, dtype=torch.int32, device="cuda")
    for flags in (0, 2):
        for _ in range(3): e.classify_slots_dev(d.data_ptr(), 64, n, flags, v.data_ptr(), st)
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(10): e.classify_slots_dev(d.data_ptr(), 64, n, flags, v.data_ptr(), st)
            t1.record(); torch.cuda.synchronize()
            best = min(best, t0.elapsed_time(t1) / 10)
        print(f"  v6={v6} flags={flags} ms={best:.4f} Gpps={n/best/1e6:.2f}", flush=True)
''' % root
for so in sys.argv[1:] * 2:                      # twice, interleaved, to see run-to-run noise
    print(so, flush=True)
    env = dict(os.environ, CLK_LIB=os.path.join(root, so))
    subprocess.run([sys.executable, "-c", code], env=env, check=False)
