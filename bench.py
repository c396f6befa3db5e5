#!/usr/bin/env python
"""bench.py — egress-verdict throughput of the clawker firewall hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own eBPF C on the host cores

A "step" is one pass of the hot path over one batch of synthetic packets.  At N = 1 the workload is
BASELINE.json configs[1] (10 M-packet synthetic pcap as 64-byte slots, 1k IP/CIDR rules); every rank of
an N-GPU run classifies its own batch of the same shape (weak scaling) and the ranks' per-GPU counters
are all-gathered once per step over NCCL.  One JSON line is printed by rank 0.

Timing: CUDA events on the launch stream, W >= 3 warm-up steps, inputs (640 MB) far larger than L2, the
max over ranks.  `value` is device-resident throughput; `e2e` is the same metric through the host-buffer
C-ABI entry (clk_classify_slots: pinned host packets -> H2D -> kernel -> D2H verdicts inside the timed
region).  `cpu_baseline` times the oracle (the reference's clawker.c compiled for the host when
oracle/_ref is present) on this box's cores — the ONLY place this file touches oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np

METRIC = "Mpackets/sec egress-verdict throughput"
WORKLOADS = {
    # IPv4, 85 % TCP / 14 % UDP / 1 % other, dst 50 % from the rule set, dport Zipf incl. random, 3 % unmanaged sources
    "c2": dict(n=10_000_000, gen=2_000_000, layout="slots", cfg="c2", mix=dict(v6_frac=0.0, nonip_frac=0.0),
               desc="10M-packet synthetic pcap (64 B slots, IPv4 85/14/1 % TCP/UDP/other), 1k IP/CIDR rules, 8 containers"),
    # the same with 2 % IPv6 and 0.5 % non-IP frames mixed in (the parity-test mix)
    "c2v6": dict(n=10_000_000, gen=2_000_000, layout="slots", cfg="c2", mix=dict(v6_frac=0.02, nonip_frac=0.005),
                 desc="10M-packet synthetic pcap (64 B slots) incl. 2 % IPv6 + 0.5 % non-IP, 1k IP/CIDR rules, 8 containers"),
    "c3": dict(n=100_000_000, gen=1_000_000, layout="varlen", cfg="c3",
               desc="100M packets incl. DNS + TLS-SNI (70/10/5/15 % header-only/DNS query/DNS answer/ClientHello), 10k-domain allowlist, name stage on"),
    # c3 plus the plaintext-HTTP Host / path-prefix stage: every 8th domain also has an http rule
    "c3h": dict(n=100_000_000, gen=1_000_000, layout="varlen", cfg="c3h", mix5=(0.64, 0.10, 0.05, 0.13, 0.08),
                desc="100M packets incl. DNS + TLS-SNI + plaintext HTTP (64/10/5/13/8 % header-only/DNS query/DNS answer/ClientHello/HTTP request), 10k-domain allowlist, name stage on"),
    "c5": dict(n=100_000_000, gen=2_000_000, layout="slots", cfg="c5:1000", desc="64 B min-size packet flood, 100M packets/GPU"),
}
# Algorithmic bytes per packet: captured bytes + 4 B verdict (+ 4 B offset in varlen mode).


# ---------------------------------------------------------------------------------------------------
def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled during the timed region (B200_PROFILING.md recipe).  NVML in-process
    when it loads (a sample every few ms, no fork — spawning nvidia-smi from a thread stalls the launching thread
    through the GIL for the length of the fork+exec, which sat inside the timed region); the nvidia-smi query of the
    recipe otherwise.  Rows: [sm MHz, max sm MHz, hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap]."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt, self._armed, self.source, self._nvml = index, [], threading.Event(), threading.Event(), None, None
        try:                                                   # NVML is attached here, before the timed region, not in the thread
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(index)
            smax = int(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = [getattr(N, "nvmlClocksEventReason" + n, None) or getattr(N, "nvmlClocksThrottleReason" + n)
                    for n in ("HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown", "SwPowerCap")]
            int(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)), int(reasons(h))          # both queries must work before we commit to NVML
            self._nvml = (N, h, smax, reasons, bits)
        except Exception:
            self._nvml = None

    def arm(self):
        """Start recording (the thread is created and parked earlier, so nothing but an Event.set() sits next to t0)."""
        self._was_armed = True
        self._armed.set()

    _was_armed = False

    def run(self):
        self._armed.wait()
        if not self._was_armed:                                # stopped without ever being armed
            return
        if self._nvml is None:
            return self._run_smi()
        N, h, smax, reasons, bits = self._nvml
        self.source = "nvml"
        self._halt.wait(0.0005)                                # first sample half a millisecond into the region: under load
        while not self._halt.is_set():
            try:
                sm, r = int(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)), int(reasons(h))
                self.rows.append([str(sm), str(smax)] + ["Active" if r & b else "Not Active" for b in bits])
            except Exception:
                pass
            self._halt.wait(0.001)
        if not self.rows:                                      # a region shorter than the first wait: one sample right after it
            try:
                sm, r = int(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)), int(reasons(h))
                self.rows.append([str(sm), str(smax)] + ["Active" if r & b else "Not Active" for b in bits])
            except Exception:
                pass

    def _run_smi(self):
        self.source = "nvidia-smi"
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self) -> dict:
        self._halt.set()
        self._armed.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows), "source": self.source}


def make_batch(workload: str, rank: int):
    """Synthetic capture of the named shape; `gen` unique packets tiled to n (HBM does not care)."""
    from clawker_b200 import capture
    spec = WORKLOADS[workload]
    w = capture.make_workload(spec["cfg"])
    if spec["layout"] == "slots":
        return w, (capture.gen_slots(w, spec["gen"], seed=2 + 1000 * rank, **spec.get("mix", {})), None), spec
    blob, offs, _ = capture.gen_varlen(w, spec["gen"], seed=3 + 1000 * rank, pool=8192, **({"mix": spec["mix5"]} if "mix5" in spec else {}))
    return w, (blob, offs), spec


def tile_host(batch, reps: int):
    """Repeat a host batch `reps` times (CPU baseline / e2e samples)."""
    pk, offs = batch
    if offs is None:
        return np.tile(pk, (reps, 1)), None
    per = int(offs[-1])
    o = np.concatenate([offs[:-1].astype(np.int64) + k * per for k in range(reps)] + [np.array([reps * per], dtype=np.int64)])
    return np.tile(pk, reps), o.astype(np.uint32)


def batch_packets(batch) -> int:
    return batch[0].shape[0] if batch[1] is None else batch[1].size - 1


# ---------------------------------------------------------------------------------------------------
def cpu_baseline(w, batch, steps: int = 1, warmup: int = 1, name_stage: bool = False) -> dict:
    """The reference's own decision code on this box's host cores (kind "reference" = /root/reference's
    clawker.c compiled for the host in oracle/_ref; "port" = the C restatement).  Bounded sample.  With the
    name stage on, the CoreDNS / Envoy / dnsbpf stages around it are the oracle's restatement in both kinds
    (third-party code, not in the reference tree)."""
    from oracle import oracle, rules_oracle
    kind = "reference" if oracle.have_ref() else "port"
    o = oracle.Oracle("ref" if kind == "reference" else "port")
    c = rules_oracle.compile_rules(w.rules)
    t = dict(w.tables)
    t.update(zones=c.zones, snis=c.snis, routes=dict(c.routes), cidrs=[] if kind == "reference" else list(c.cidrs))
    o.load(t)
    cores = os.cpu_count() or 1
    npk = batch_packets(batch)
    reps = max(1, (250_000 * cores) // npk)                         # ~0.25 M packets per thread: 10-30 CPU-seconds per step
    pk, offs = tile_host(batch, reps) if reps > 1 else batch
    flags = 1 if name_stage else 0
    n_s = batch_packets((pk, offs))
    verd = np.zeros(n_s, dtype=np.uint32)
    for _ in range(max(1, warmup)):                                 # untimed: warms the tables and the page cache
        o.classify_timed(pk, offs, flags=flags, nthreads=cores, verd=verd)
    # only the parallel section is timed: the threads stand for the CPUs the hooks run on and already exist, each has
    # its own metrics map (BPF_MAP_TYPE_PERCPU_HASH) that nothing merges on the hot path
    dt = sum(o.classify_timed(pk, offs, flags=flags, nthreads=cores, verd=verd) for _ in range(max(1, steps))) / max(1, steps)
    sample = (f"{n_s} packets of the same capture per step, {cores} threads, per-thread metrics maps (PERCPU, not merged), "
              "thread start-up excluded"
              + ("; IP/CIDR extension rules are not evaluated by the reference (SURVEY R4)" if kind == "reference" else ""))
    return {"value": round(n_s / dt / 1e6, 2), "unit": "Mpackets/s", "cores": cores, "kind": kind, "sample": sample,
            "ms_per_step": round(dt * 1e3, 3), "packets_per_step": n_s}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w, batch, spec = make_batch(args.workload, 0)
    t0 = time.perf_counter()
    base = cpu_baseline(w, batch, steps=max(1, args.steps), warmup=max(1, args.warmup), name_stage=spec["layout"] == "varlen")
    ms, pps = base.pop("ms_per_step"), base.pop("packets_per_step")
    line = {
        "metric": METRIC, "impl": "reference", "value": base["value"], "unit": "Mpackets/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": spec["desc"], "packets_per_step": pps, "egress_rules": len(w.rules),
                   "layout": "varlen (16 B aligned frames + u32 offsets)" if spec["layout"] == "varlen" else "64 B slots",
                   "parallelism": f"host threads x{base['cores']} (the reference's CPU path; rank 0 only)"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "Mpackets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": round(time.perf_counter() - t0, 2),
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def run_gpu(args) -> None:
    import torch
    import torch.distributed as dist
    from clawker_b200 import abi, engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the CUDA library is the product, there is no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    w, batch, spec = make_batch(args.workload, rank)
    pk, offs = batch
    varlen = offs is not None
    gen = batch_packets(batch)
    n = spec["n"] if args.packets <= 0 else args.packets
    reps = (n + gen - 1) // gen
    n = reps * gen if varlen else n                                # varlen batches are tiled whole
    dev = torch.device("cuda", local)
    if varlen:
        d_pk = torch.from_numpy(pk).to(dev).repeat(reps)
        per = int(offs[-1])
        o64 = torch.from_numpy(offs.astype(np.int64)).to(dev)
        d_offs = torch.cat([o64[:-1] + k * per for k in range(reps)] + [o64[-1:] + (reps - 1) * per]).to(torch.int32)
        total_bytes = int(d_pk.numel())
        if (total_bytes >> 4) >= (1 << 32):
            raise SystemExit("batch exceeds the 64 GiB reach of 16-byte offsets")
        bytes_per_packet = total_bytes / n + 8
    else:
        d_unique = torch.from_numpy(pk).to(dev)
        d_pk = d_unique.repeat(reps, 1)[:n].contiguous()
        del d_unique
        d_offs = None
        bytes_per_packet = 68.0
    d_v = torch.empty(n, dtype=torch.int32, device=dev)
    flags = 1 if varlen else 0                                     # CLK_F_NAME_STAGE on the varlen workload

    # metrics_map max_entries: 8x the reference default (common.h:173) so every key of the synthetic capture is counted;
    # keys beyond max_entries are dropped like the reference BPF_NOEXIST insert into a full map (--metrics-capacity)
    # multi-GPU: the classify kernels leave a few SMs to the side stream so the per-batch collective never delays them
    reserved = args.reserved_sms if args.reserved_sms >= 0 else (4 if world > 1 else 0)
    e = engine.Engine(device=local, max_metrics=args.metrics_capacity, max_dns_cache=1 << 21, max_routes=1 << 16, reserved_sms=reserved)
    e.load_tables(w.tables)
    e.compile_rules(w.rules)
    e.publish()
    rows = e.counters_block_rows()
    d_block = torch.empty(rows * 24, dtype=torch.uint8, device=dev)
    d_gather = torch.empty(world * rows * 24, dtype=torch.uint8, device=dev) if world > 1 else None
    trow = e.totals_rows()
    d_tot = torch.zeros(trow * 32, dtype=torch.uint8, device=dev)
    d_tot_all = torch.zeros(world * trow * 32, dtype=torch.uint8, device=dev) if world > 1 else None

    ts = torch.cuda.Stream(device=dev)
    side = torch.cuda.Stream(device=dev)                           # the collective overlaps the next batch's kernel
    torch.cuda.set_stream(ts)
    st = ts.cuda_stream

    def step(ev_pair=None):
        if ev_pair:
            ev_pair[0].record(ts)
        if varlen:
            e.classify_varlen_dev(d_pk.data_ptr(), d_offs.data_ptr(), n, flags, d_v.data_ptr(), st)
        else:
            e.classify_slots_dev(d_pk.data_ptr(), 64, n, flags, d_v.data_ptr(), st)
        if ev_pair:
            ev_pair[1].record(ts)
        if world > 1:                                              # per batch: per-container totals of every GPU over NVLink
            side.wait_stream(ts)
            with torch.cuda.stream(side):
                e.totals_export(d_tot.data_ptr(), side.cuda_stream)
                dist.all_gather_into_tensor(d_tot_all, d_tot)

    def merge_on_read():                                           # per-key merge of the metrics_map tables: once per read
        with torch.cuda.stream(side):
            e.counters_export(d_block.data_ptr(), side.cuda_stream)
            dist.all_gather_into_tensor(d_gather, d_block)
            e.counters_import(d_gather.data_ptr(), world, side.cuda_stream)

    sampler = ClockSampler(local) if rank == 0 else None          # attached and parked before the barrier: rank 0 must not start late
    if sampler:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        step()
    if world > 1:
        side.wait_stream(ts)
        merge_on_read()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = e.stats().kernel_launches
    kern_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if sampler:
        sampler.arm()
    t0.record(ts)
    for k in range(args.steps):
        step(kern_events[k])
    ts.wait_stream(side)
    t1.record(ts)
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None                   # samples of the timed region only
    if world > 1:
        dist.barrier()
    ms_total = t0.elapsed_time(t1)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kern_events]))
    launches = e.stats().kernel_launches - launches0               # our kernels inside the timed region (classify [+ totals])
    merge_ms = None
    if world > 1:                                                  # what one reader of the merged metrics_map costs (not a per-batch step)
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record(side)
        merge_on_read()
        m1.record(side)
        torch.cuda.synchronize()
        merge_ms = m0.elapsed_time(m1)
        tot = e.totals_merge(d_tot_all.data_ptr(), world)
        mg = e.counters(merged=True)
        for a in range(3):                                         # the two forms of the sum must agree
            assert int(tot["count"][:, a].sum()) == int(mg["count"][mg["action"] == a].sum()), "per-container totals != merged metrics_map"

    # ---- end to end: host buffers through the C-ABI (H2D + kernel + D2H inside the timed region) ----
    e2e_reps = max(1, min(n, 10_000_000) // gen)
    hp, ho = tile_host(batch, e2e_reps) if e2e_reps > 1 else batch
    n_e2e = batch_packets((hp, ho))
    h_pk = engine.pinned_empty(hp.size).reshape(hp.shape)
    h_pk[...] = hp
    h_v = engine.pinned_empty(n_e2e * 4).view(np.uint32)
    if varlen:                                                     # the offsets ride in the pinned ring with the frames
        h_off = engine.pinned_empty(ho.size * 4).view(np.uint32)
        h_off[...] = ho
        ho = h_off
    h2d_bytes = int(hp.size) + (int(ho.size) * 4 if varlen else 0)

    def e2e_call():
        if varlen:
            e.classify_varlen(h_pk, ho, flags, out=h_v)
        else:
            e.classify_slots(h_pk, 64, flags, out=h_v)
    for _ in range(2):
        e2e_call()
    if world > 1:
        dist.barrier()
    e2e_steps = max(2, min(args.steps, 5))
    tw0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_call()
    e2e_ms = (time.perf_counter() - tw0) * 1e3 / e2e_steps

    # ---- max over ranks ------------------------------------------------------------------------------
    if world > 1:
        tt = torch.tensor([ms_total, kern_ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, kern_ms, e2e_ms = (float(x) for x in tt.tolist())
    if rank == 0:
        peak, how = peaks()
        ms_per_step = ms_total / args.steps
        value = world * n / ms_per_step / 1e3                     # Mpackets/s, whole job
        achieved = n * bytes_per_packet / (kern_ms * 1e-3) / 1e9
        prof = ROOT / "profiles" / ("r01_traffic.json" if not varlen else "r01_varlen_traffic.json")
        traffic = None
        if prof.exists():                                          # ncu dram__bytes_{read,write}.sum of one launch, scaled to this launch's packets
            pj = json.loads(prof.read_text())
            traffic = round(pj["dram_bytes_per_launch"] * n / pj["packets_per_launch"])
        base = None
        if world == 1 and not args.no_cpu_baseline:
            base = cpu_baseline(w, batch, steps=2, warmup=1, name_stage=varlen)
            base.pop("ms_per_step"), base.pop("packets_per_step")
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "Mpackets/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": spec["desc"], "packets_per_gpu_per_step": n, "egress_rules": len(w.rules), "table_rows": {"dns_cache": e.map_count(2), "route_map": e.map_count(3), "cidr": e.map_count(7), "containers": e.map_count(0)}, "layout": "varlen (16 B aligned frames + u32 offsets)" if varlen else "64 B slots",
                       "l2_policy": f"inputs_larger_than_l2 ({n * bytes_per_packet / 1e6:.0f} MB per step vs 126 MB L2)",
                       "parallelism": (f"packet shards x{world}, tables replicated, no data-path collective; per step 1 NCCL all-gather of per-container totals "
                                       f"({trow * 32} B per GPU, side stream, {reserved} SMs reserved); per-key metrics_map merge on read, outside the timed region "
                                       f"({merge_ms:.3f} ms per read)") if world > 1 else "single GPU"},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": traffic, "peak_source": how, "kernel": "clk::k_varlen_fast" if varlen else "clk::k_slots_fast", "kernel_ms": round(kern_ms, 4),
                         "algorithmic_bytes_per_packet": round(bytes_per_packet, 2)},
            "e2e": {"value": round(world * n_e2e / e2e_ms / 1e3, 1), "unit": "Mpackets/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": n_e2e * 4, "packets_per_step": n_e2e, "ms_per_step": round(e2e_ms, 3),
                    "path": ("clk_classify_varlen" if varlen else "clk_classify_slots") + " (pinned host buffers, 3-stream chunked H2D/kernel/D2H)"},
            "gpu_launches": int(launches), "clocks": clocks,
            "metrics": {"max_entries": args.metrics_capacity, "rows": int(e.counters().size), "dropped_increments": int(e.stats().metrics_dropped)},
        }
        if base:
            line["cpu_baseline"] = base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--packets", type=int, default=0, help="override packets per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--reserved-sms", type=int, default=-1, help="SMs the classify kernels leave to the side-stream collective (default: 4 when N > 1, else 0)")
    ap.add_argument("--metrics-capacity", type=int, default=131072, help="metrics_map max_entries per GPU (reference: 16384; the synthetic capture produces ~76k distinct keys)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
