"""Where does the multi-GPU step time go?  torchrun --nproc-per-node 2 scripts/probe_multi.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from clawker_b200 import capture, engine
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 10_000_000
w = capture.make_workload("c2"); pk = capture.gen_slots(w, 2_000_000, seed=2 + rank)
d = torch.from_numpy(pk).cuda().repeat(5, 1).contiguous(); v = torch.empty(n, dtype=torch.int32, device="cuda")
e = engine.Engine(device=local, max_metrics=131072); e.load_tables(w.tables); e.compile_rules(w.rules); e.publish()
rows = e.counters_block_rows(); blk = torch.empty(rows * 24, dtype=torch.uint8, device="cuda"); g = torch.empty(world * rows * 24, dtype=torch.uint8, device="cuda")
ts = torch.cuda.Stream(); side = torch.cuda.Stream(); torch.cuda.set_stream(ts)
def classify(): e.classify_slots_dev(d.data_ptr(), 64, n, 0, v.data_ptr(), ts.cuda_stream)
def coll(st):
    e.counters_export(blk.data_ptr(), st.cuda_stream)
    with torch.cuda.stream(st): dist.all_gather_into_tensor(g, blk)
    e.counters_import(g.data_ptr(), world, st.cuda_stream)
def run(name, fn, K=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    h0 = time.perf_counter(); t0.record(ts)
    for _ in range(K): fn()
    ts.wait_stream(side); t1.record(ts); h1 = time.perf_counter(); torch.cuda.synchronize()
    if rank == 0: print(f"{name:28s} gpu {t0.elapsed_time(t1)/K:.3f} ms/step   host enqueue {1e3*(h1-h0)/K:.3f} ms/step", flush=True)
run("classify only", classify)
run("collective only (same stream)", lambda: coll(ts))
def serial(): classify(); coll(ts)
run("classify + coll serial", serial)
def overlap(): classify(); side.wait_stream(ts); coll(side)
run("classify + coll overlapped", overlap)
def export_only(): e.counters_export(blk.data_ptr(), ts.cuda_stream)
run("export only", export_only)
def ag_only(): dist.all_gather_into_tensor(g, blk)
run("all_gather only", ag_only)
def import_only(): e.counters_import(g.data_ptr(), world, ts.cuda_stream)
run("import only", import_only)
dist.destroy_process_group()
