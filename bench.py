#!/usr/bin/env python
"""bench.py — egress-verdict throughput of the clawker firewall hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own eBPF C on the host cores

A "step" is one pass of the hot path over one batch of synthetic packets.  The default workload is the one the
north-star target is quoted on, BASELINE.json configs[2] ("c3": 100 M packets incl. DNS + TLS-SNI, 10k-domain
allowlist, name stage on, variable-length frames); configs[1] ("c2": 10 M 64-byte slots, 1k IP/CIDR rules) is
measured in the same run and reported inside the same JSON line under "also".  Every rank of an N-GPU run
classifies its own batch of the same shape (weak scaling); the ranks' per-container counter totals are
all-gathered once per step over NCCL.  One JSON line is printed by rank 0.

Timing: CUDA events on the launch stream, W >= 3 warm-up steps, inputs far larger than L2, the max over ranks.
`value` is device-resident throughput of exactly K steps; `sustained` repeats the step for >= --seconds of GPU
time; `e2e` is the same metric through the host-buffer C-ABI entry (pinned host packets -> H2D -> kernel -> D2H
verdicts inside the timed region); `parity` compares the verdict words and the metrics_map of the benched
batch with the CPU oracle (outside the timed region).  `cpu_baseline` / `--impl reference` time the oracle (the
reference's clawker.c compiled for the host when oracle/_ref is present) on this box's cores — parity and those
two legs are the ONLY places this file touches oracle/.
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
    # BASELINE.json configs[3]: C2 u C3 mix, 100k rules, 125 M packets per GPU (1 B over 8 GPUs)
    "c4": dict(n=125_000_000, gen=1_000_000, layout="varlen", cfg="c4", mix=(0.85, 0.05, 0.03, 0.07),
               desc="1B mixed TCP/UDP/DNS packets over 8 GPUs (125M per GPU; 85/5/3/7 % header-only/DNS query/DNS answer/ClientHello), 100k rules, 64 containers, name stage on"),
    # BASELINE.json configs[4]: 64 B flood against a rule table of the given size ("c5:<entries>", --sweep runs 2^4..2^20)
    "c5": dict(n=100_000_000, gen=2_000_000, layout="slots", cfg="c5:1024", mix=dict(v6_frac=0.0, nonip_frac=0.0),
               desc="64 B min-size packet flood, 100M packets/GPU, rule table of 1024 entries"),
}
SWEEP_SIZES = [1 << k for k in range(4, 21, 2)]
# Algorithmic bytes per packet: captured bytes + 4 B verdict (+ 4 B offset in varlen mode).


def workload_spec(name: str) -> dict:
    """WORKLOADS entry; "c5:<entries>" selects the rule-table size of the flood workload."""
    if name.startswith("c5:"):
        k = int(name.split(":")[1])
        s = dict(WORKLOADS["c5"])
        s["cfg"] = f"c5:{k}"
        s["desc"] = f"64 B min-size packet flood, 100M packets/GPU, rule table of {k} entries"
        return s
    return WORKLOADS[name]


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

    def __init__(self, index: int, period: float = 0.001):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt, self._armed, self.source, self._nvml = index, [], threading.Event(), threading.Event(), None, None
        self.period = period
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
            self._halt.wait(self.period)
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


def bind_to_gpu_numa(local: int) -> dict:
    """Run this rank (and first-touch its pinned staging buffers) on the NUMA node its GPU hangs off: at N = 8 every rank pulls
    ~54 GB/s out of host memory, and buffers on the far socket halve that (round-1 e2e scaling 0.72 at N = 8)."""
    info = {"numa_node": None}
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{getattr(pr, 'pci_domain_id', 0):04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(Path(f"/sys/bus/pci/devices/{bdf}/numa_node").read_text())
        info["pci"] = bdf
        if node < 0:
            return info
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(numa_node=node, cpus=len(allowed))
    except Exception as ex:                                     # no sysfs entry, no permission: run unbound
        info["error"] = str(ex)[:80]
    return info


def make_batch(workload: str, rank: int, untiled: bool = False):
    """Synthetic capture of the named shape: `gen` unique packets (tiled to n on the device; HBM does not care —
    the tiling is stated in config.synthetic and --untiled generates every packet of the batch)."""
    from clawker_b200 import capture
    spec = dict(workload_spec(workload))
    if untiled:
        spec["gen"] = spec["n"]
    w = capture.make_workload(spec["cfg"])
    if spec["layout"] == "slots":
        return w, (capture.gen_slots(w, spec["gen"], seed=2 + 1000 * rank, **spec.get("mix", {})), None), spec
    kw = {}
    if "mix5" in spec:
        kw["mix"] = spec["mix5"]
    elif "mix" in spec:
        kw["mix"] = spec["mix"]
    blob, offs, _ = capture.gen_varlen(w, spec["gen"], seed=3 + 1000 * rank, pool=8192, **kw)
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


def table_rows(w) -> dict:
    """Rows behind the rule count, from the workload definition alone (so both arms print the same numbers)."""
    return {"dns_cache": len(w.tables["dns_cache"]), "route_map": int(w.meta["n_routes"]), "cidr": int(w.meta["n_cidr"]),
            "containers": len(w.tables["containers"]), "domains": int(w.meta["n_domains"])}


def config_of(spec: dict, w, n: int, gen: int, world: int) -> dict:
    """The `config` object — identical for the CUDA arm and the reference arm of the same command line."""
    varlen = spec["layout"] == "varlen"
    return {"workload": spec["desc"], "packets_per_gpu_per_step": n, "egress_rules": len(w.rules), "table_rows": table_rows(w),
            "layout": "varlen (16 B aligned frames + u32 offsets)" if varlen else "64 B slots",
            "synthetic": f"{gen} unique seeded packets per GPU tiled x{max(1, n // gen)} to the batch size",
            "l2_policy": "inputs_larger_than_l2 (>= 640 MB per step vs 126 MB L2)",
            "parallelism": f"packet shards x{world} (weak), tables replicated, no data-path collective; per step 1 all-gather of per-container counter totals" if world > 1 else "single GPU"}


# ---------------------------------------------------------------------------------------------------
# CPU side: the oracle as checker (parity) and as baseline (cpu_baseline / --impl reference)
# ---------------------------------------------------------------------------------------------------
def _oracle_for(w, name_stage: bool, want_cidr: bool):
    """Loaded oracle + its kind.  The reference object ignores IP/CIDR rules, so a workload with CIDR
    rules is CHECKED against the port (which evaluates them, A12) and TIMED on the reference object."""
    from oracle import oracle, rules_oracle
    c = rules_oracle.compile_rules(w.rules)
    use_ref = oracle.have_ref() and not (want_cidr and c.cidrs)
    o = oracle.Oracle("ref" if use_ref else "port")
    t = dict(w.tables)
    t.update(zones=c.zones, snis=c.snis, routes=dict(c.routes), cidrs=[] if use_ref else list(c.cidrs), http=list(getattr(c, "http", []) or []))
    o.load(t)
    return o, ("reference" if use_ref else "port")


def cpu_quota() -> float | None:
    """CPUs the container may use per the cgroup v2 quota (cpu.max "quota period"), None when unlimited / unknown."""
    try:
        q, per = Path("/sys/fs/cgroup/cpu.max").read_text().split()[:2]
        return None if q == "max" else float(q) / float(per)
    except Exception:
        return None


def host_threads() -> int:
    """Threads the CPU arm runs: one per CPU of the affinity mask, but no more than the cgroup CPU quota allows — threads
    beyond the quota are only throttled (the GPU boxes show 128 CPUs under a 16-CPU quota)."""
    try:
        n = max(1, len(os.sched_getaffinity(0)))
    except Exception:
        n = os.cpu_count() or 1
    q = cpu_quota()
    if q is not None:
        n = max(1, min(n, int(q + 0.5)))
    return n


def cpu_baseline(w, batch, samples: int = 5, warmup: int = 1, name_stage: bool = False, target_s: float = 2.0) -> dict:
    """The reference's own decision code on this box's host cores (kind "reference" = /root/reference's clawker.c
    compiled for the host in oracle/_ref; "port" = the C restatement).  With the name stage on, the CoreDNS / Envoy /
    dnsbpf stages around it are the oracle's restatement in both kinds (third-party code, not in the reference tree).

    One sample = every thread (pinned, one per allowed CPU) classifies its private copy of ~0.25 M packets of the same
    capture `repeat` times, `repeat` calibrated so a sample lasts about `target_s` seconds; the value is the MEDIAN of
    `samples` samples (min and max reported)."""
    o, kind = _oracle_for(w, name_stage, want_cidr=False)
    cores = host_threads()
    npk = batch_packets(batch)
    reps = max(1, (250_000 * cores) // npk)
    pk, offs = tile_host(batch, reps) if reps > 1 else batch
    flags = 1 if name_stage else 0
    n_s = batch_packets((pk, offs))
    verd = np.zeros(n_s, dtype=np.uint32)
    o.classify_timed2(pk, offs, flags=flags, nthreads=cores, repeat=1, verd=verd)                # cold pass: page-in, tables
    dt1 = o.classify_timed2(pk, offs, flags=flags, nthreads=cores, repeat=1, verd=verd)          # calibration pass
    repeat = max(1, int(round(target_s / max(dt1, 1e-4))))
    for _ in range(max(0, warmup)):
        o.classify_timed2(pk, offs, flags=flags, nthreads=cores, repeat=repeat, verd=verd)
    dts = sorted(o.classify_timed2(pk, offs, flags=flags, nthreads=cores, repeat=repeat, verd=verd) for _ in range(max(1, samples)))
    med = dts[len(dts) // 2]
    per = n_s * repeat
    sample = (f"{n_s} packets of the same capture x{repeat} passes per sample ({per} verdicts, ~{med:.2f} s), {cores} threads pinned one per CPU, "
              f"thread-private copies of the input, per-thread metrics maps (PERCPU, not merged), thread start-up excluded; median of {len(dts)} samples"
              + ("; IP/CIDR extension rules are not evaluated by the reference (SURVEY R4)" if kind == "reference" else ""))
    return {"value": round(per / med / 1e6, 2), "unit": "Mpackets/s", "cores": cores, "kind": kind, "sample": sample,
            "cpus_visible": os.cpu_count(), "cpu_quota": cpu_quota(),
            "min": round(per / dts[-1] / 1e6, 2), "max": round(per / dts[0] / 1e6, 2), "samples": len(dts),
            "ms_per_step": round(med * 1e3, 3), "packets_per_step": per}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    w, batch, spec = make_batch(args.workload, 0)
    t0 = time.perf_counter()
    base = cpu_baseline(w, batch, samples=max(1, args.steps), warmup=max(0, args.warmup), name_stage=spec["layout"] == "varlen",
                        target_s=args.ref_sample_seconds)
    ms = base.pop("ms_per_step")
    base.pop("packets_per_step")
    gen = batch_packets(batch)
    n = spec["n"] if args.packets <= 0 else args.packets
    n = ((n + gen - 1) // gen) * gen if spec["layout"] == "varlen" else n
    line = {
        "metric": METRIC, "impl": "reference", "value": base["value"], "unit": "Mpackets/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": config_of(spec, w, n, gen, max(world, args.gpus)),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "Mpackets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": round(time.perf_counter() - t0, 2),
    }
    print(json.dumps(line), flush=True)


def check_parity(e, w, batch, d_v, n: int, gen: int, varlen: bool, flags: int, run_step, threads: int) -> dict:
    """Bit-exact check of the BENCHED batch (outside the timed region): the verdict words of the first and the last
    whole copy of the unique packets, and — after a counters reset and one more pass — this GPU's metrics_map against
    the oracle's (counts x the tiling factor)."""
    import torch
    from oracle import oracle
    o, kind = _oracle_for(w, varlen, want_cidr=True)
    pk, offs = batch
    want, wm = o.classify(pk, offs, flags=flags, nthreads=threads)
    reps = n // gen
    blocks = sorted({0, max(0, reps - 1)})
    checked = mism = 0
    for b in blocks:
        got = d_v[b * gen:(b + 1) * gen].cpu().numpy().view(np.uint32)
        k = min(got.size, want.size)
        mism += int((got[:k] != want[:k]).sum())
        checked += k
    out = {"checked": checked, "mismatches": mism, "rate": round(1.0 - mism / max(1, checked), 9), "oracle": "ref" if kind == "reference" else "port"}
    if reps > 1:                                                   # every other copy of the unique packets, on the device: equal to the first copy word for word
        whole = d_v[:reps * gen].view(reps, gen)
        out["copies_checked"] = reps                               # the whole benched batch: reps x gen verdict words
        out["copy_mismatches"] = int((whole != whole[0:1]).sum().item())
        out["rate"] = round(1.0 - (mism + out["copy_mismatches"]) / float(reps * gen), 9)
    if n == reps * gen:
        e.counters_reset()
        run_step()
        torch.cuda.synchronize()
        gm, dropped = e.counters(), int(e.stats().metrics_dropped)
        if dropped == 0:
            a, b = oracle.sort_metrics(gm), oracle.sort_metrics(wm)
            same_keys = a.size == b.size and all(np.array_equal(a[f], b[f]) for f in ("cgroup_id", "domain_hash", "dst_port", "action"))
            bad = int((a["count"] != b["count"] * reps).sum()) if same_keys else max(a.size, b.size)
            out["counters"] = {"rows": int(a.size), "oracle_rows": int(b.size), "mismatches": bad}
        else:
            out["counters"] = {"rows": int(gm.size), "skipped": f"metrics_map reached max_entries ({dropped} increments dropped, as BPF_NOEXIST into a full map)"}
        out["totals_oracle"] = [int(wm["count"][wm["action"] == a].sum()) * reps for a in range(3)]
    return out


# ---------------------------------------------------------------------------------------------------
def measure(args, workload: str, full: bool) -> dict:
    """One workload on this rank's GPU: device-resident K steps (+ sustained), parity, e2e.  Returns the pieces of the
    JSON line (rank 0) — `full` adds the sustained leg and the CPU baseline."""
    import torch
    import torch.distributed as dist
    from clawker_b200 import abi, engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    w, batch, spec = make_batch(workload, rank, untiled=args.untiled)
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
        del o64
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
    flags = abi.F_NAME_STAGE if varlen else 0
    if os.environ.get("CLK_BENCH_FLAGS"):                          # experiments only (e.g. 0: the header ladder alone on the varlen layout)
        flags = int(os.environ["CLK_BENCH_FLAGS"])

    # metrics_map max_entries: 8x the reference default (common.h:173) so every key of the synthetic capture is counted;
    # keys beyond max_entries are dropped like the reference BPF_NOEXIST insert into a full map (--metrics-capacity)
    # multi-GPU: the classify kernels leave a few SMs to the side stream so the per-batch collective never delays them
    reserved = args.reserved_sms if args.reserved_sms >= 0 else (4 if world > 1 and not varlen else 0)   # ms-long varlen steps do not need them
    e = engine.Engine(device=local, max_metrics=args.metrics_capacity, max_dns_cache=1 << 21, max_routes=1 << 16, max_containers=max(256, len(w.tables["containers"])),
                      reserved_sms=reserved)
    e.load_tables(w.tables)
    e.compile_rules(w.rules)
    e.publish()
    if world > 1:                                                  # the communicator lives behind the C ABI (clk_comm_*): NCCL over NVLink
        ids = [engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        e.comm_init(ids[0], rank, world)
    trow = e.totals_rows()

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
            e.allgather_counters(side.cuda_stream)                 # clk_allgather_counters: totals export + ncclAllGather, enqueue only

    def merge_on_read():                                           # per-key merge of the metrics_map tables: once per read
        e.allgather_metrics(side.cuda_stream)                      # clk_allgather_metrics: export + ncclAllGather + import

    sampler = ClockSampler(local) if rank == 0 else None          # attached and parked before the barrier: rank 0 must not start late
    if sampler:
        sampler.start()
    warm = max(3, args.warmup)
    for _ in range(warm):
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

    # ---- sustained: the same step back to back for >= --seconds of GPU time (clocks and power under real load) -----
    sustained = None
    if full and args.seconds > 0:
        s2 = ClockSampler(local, period=0.02) if rank == 0 else None
        if s2:
            s2.start()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        est = max(ms_total / max(1, args.steps), 1e-3)
        total_steps = max(args.steps, int(args.seconds * 1e3 / est) + 1)
        if world > 1:                                              # every rank must run the same number of steps (one collective each)
            tsn = torch.tensor([total_steps], dtype=torch.int64, device=dev)
            dist.all_reduce(tsn, op=dist.ReduceOp.MAX)
            total_steps = int(tsn.item())
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if s2:
            s2.arm()
        u0.record(ts)
        for _ in range(total_steps):
            step()
        ts.wait_stream(side)
        u1.record(ts)
        torch.cuda.synchronize()
        c2 = s2.stop() if s2 else None
        sus_ms = u0.elapsed_time(u1)
        if world > 1:
            tt = torch.tensor([sus_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            sus_ms = float(tt.item())
        sustained = {"seconds": round(sus_ms / 1e3, 3), "steps": total_steps, "ms_per_step": round(sus_ms / total_steps, 4),
                     "value": round(world * n * total_steps / sus_ms / 1e3, 1), "unit": "Mpackets/s", "clocks": c2}

    merge_ms = None
    tot = None
    if world > 1:                                                  # what one reader of the merged metrics_map costs (not a per-batch step)
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record(side)
        merge_on_read()
        m1.record(side)
        torch.cuda.synchronize()
        merge_ms = m0.elapsed_time(m1)
        tot = e.totals_global()
        mg = e.counters(merged=True)
        for a in range(3):                                         # the two forms of the sum must agree
            assert int(tot["count"][:, a].sum()) == int(mg["count"][mg["action"] == a].sum()), "per-container totals != merged metrics_map"

    # ---- parity of the benched batch (outside the timed region) -------------------------------------------------------
    parity = None
    if not args.no_parity:
        parity = check_parity(e, w, batch, d_v, n, gen, varlen, flags, step, max(1, host_threads() // max(1, world)))
        if world > 1:
            side.wait_stream(ts)
            torch.cuda.synchronize()
            tot1 = e.totals_global()                               # totals gathered by the parity pass: one batch per rank since the reset
            mine = torch.tensor([parity["checked"], parity["mismatches"], parity.get("counters", {}).get("mismatches", 0) or 0]
                                + parity.get("totals_oracle", [0, 0, 0]), dtype=torch.int64, device=dev)
            dist.all_reduce(mine, op=dist.ReduceOp.SUM)
            v = [int(x) for x in mine.tolist()]
            parity.update(checked=v[0], mismatches=v[1], rate=round(1.0 - v[1] / max(1, v[0]), 9))
            if "counters" in parity and "mismatches" in parity["counters"]:
                parity["counters"]["mismatches"] = v[2]
            got_tot = [int(tot1["count"][:, a].sum()) for a in range(3)]
            parity["merged_totals"] = {"gathered": got_tot, "oracle_sum_over_ranks": v[3:6], "equal": got_tot == v[3:6]}
        parity.pop("totals_oracle", None)

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
    e2e_parity = None
    if parity is not None:                                         # the host-buffer path writes the same words
        from oracle import oracle as _o                            # noqa: F401  (checker only)
        o, _k = _oracle_for(w, varlen, want_cidr=True)
        want, _ = o.classify(pk, offs, flags=flags, nthreads=max(1, host_threads() // max(1, world)), metrics=False)
        e2e_parity = int((np.asarray(h_v[:gen]) != want[:gen]).sum()) + int((np.asarray(h_v[n_e2e - gen:]) != want[:gen]).sum())

    # ---- end to end, compact ingest (slot workloads): 16-byte records over the host link instead of 64-byte slots ----------
    e2e_compact = None
    if not varlen:
        tp0 = time.perf_counter()
        recs, aux = engine.pack_hdr16(hp, pinned=True)             # the capture front end's job (clk_pack_hdr16, one thread here); not in the timed region
        pack_s = time.perf_counter() - tp0
        for _ in range(2):
            e.classify_hdr16(recs, aux, flags, out=h_v)
        if world > 1:
            dist.barrier()
        tc0 = time.perf_counter()
        for _ in range(e2e_steps):
            e.classify_hdr16(recs, aux, flags, out=h_v)
        c_ms = (time.perf_counter() - tc0) * 1e3 / e2e_steps
        c_bad = None
        if parity is not None:
            c_bad = int((np.asarray(h_v[:gen]) != want[:gen]).sum()) + int((np.asarray(h_v[n_e2e - gen:]) != want[:gen]).sum())
        e2e_compact = {"ms": c_ms, "h2d": int(recs.nbytes + aux.nbytes), "bad": c_bad, "pack_mpps_one_thread": round(n_e2e / pack_s / 1e6, 1), "aux_slots": int(aux.size // 64)}

    # ---- max over ranks ------------------------------------------------------------------------------
    if world > 1:
        tt = torch.tensor([ms_total, kern_ms, e2e_ms, e2e_compact["ms"] if e2e_compact else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, kern_ms, e2e_ms, cms = (float(x) for x in tt.tolist())
        if e2e_compact:
            e2e_compact["ms"] = cms
        if e2e_parity is not None:
            ep = torch.tensor([e2e_parity], dtype=torch.int64, device=dev)
            dist.all_reduce(ep, op=dist.ReduceOp.SUM)
            e2e_parity = int(ep.item())
    out = None
    if rank == 0:
        peak, how = peaks()
        ms_per_step = ms_total / args.steps
        value = world * n / ms_per_step / 1e3                     # Mpackets/s, whole job
        achieved = n * bytes_per_packet / (kern_ms * 1e-3) / 1e9
        traffic = None
        for cand in (("r02_varlen_traffic.json", "r01_varlen_traffic.json") if varlen else ("r02_traffic.json", "r01_traffic.json")):
            prof = ROOT / "profiles" / cand
            if prof.exists():                                      # ncu dram__bytes_{read,write}.sum of one launch, scaled to this launch's packets
                pj = json.loads(prof.read_text())
                traffic = round(pj["dram_bytes_per_launch"] * n / pj["packets_per_launch"])
                break
        cfg = config_of(spec, w, n, gen, world)
        if world > 1:
            cfg["collective"] = (f"clk_allgather_counters (NCCL behind the C ABI): {trow * 32} B per GPU per step on a side stream, {reserved} SMs reserved; "
                                 f"per-key metrics_map merge on read (clk_allgather_metrics), outside the timed region ({merge_ms:.3f} ms per read)")
        out = {
            "value": round(value, 1), "unit": "Mpackets/s", "ms_per_step": round(ms_per_step, 4), "config": cfg,
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": traffic, "peak_source": how, "kernel": ("clk varlen pipeline: k_varlen_head + k_generic_list + k_name_list<dns query | dns/tcp | client hello | dns answer | http> per slice (timed as one)"
                                    if varlen else "clk::k_slots_fast"), "kernel_ms": round(kern_ms, 4),
                         "algorithmic_bytes_per_packet": round(bytes_per_packet, 2)},
            "e2e": {"value": round(world * n_e2e / e2e_ms / 1e3, 1), "unit": "Mpackets/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": n_e2e * 4, "packets_per_step": n_e2e, "ms_per_step": round(e2e_ms, 3),
                    "bytes_per_packet_over_pcie": round((h2d_bytes + n_e2e * 4) / n_e2e, 1), "h2d_GBps_per_gpu": round(h2d_bytes / e2e_ms / 1e6, 1), "verdict_mismatches": e2e_parity,
                    "path": ("clk_classify_varlen" if varlen else "clk_classify_slots") + " (pinned host buffers, 3-stream chunked H2D/kernel/D2H)"},
            "gpu_launches": int(launches), "clocks": clocks, "parity": parity,
            "metrics": {"max_entries": args.metrics_capacity, "rows": int(e.counters().size), "dropped_increments": int(e.stats().metrics_dropped)},
        }
        if e2e_compact:
            out["e2e_compact"] = {"value": round(world * n_e2e / e2e_compact["ms"] / 1e3, 1), "unit": "Mpackets/s", "h2d_bytes_per_step": e2e_compact["h2d"],
                                  "d2h_bytes_per_step": n_e2e * 4, "packets_per_step": n_e2e, "ms_per_step": round(e2e_compact["ms"], 3),
                                  "bytes_per_packet_over_pcie": round((e2e_compact["h2d"] + n_e2e * 4) / n_e2e, 1), "h2d_GBps_per_gpu": round(e2e_compact["h2d"] / e2e_compact["ms"] / 1e6, 1), "verdict_mismatches": e2e_compact["bad"],
                                  "aux_slots": e2e_compact["aux_slots"], "pack_mpps_one_thread": e2e_compact["pack_mpps_one_thread"],
                                  "path": "clk_classify_hdr16: 16-byte ingest records (clk_pack_hdr16, packed by the capture front end outside the timed region) + whole frames for what a record cannot express"}
        if sustained:
            out["sustained"] = sustained
        if full and world == 1 and not args.no_cpu_baseline:
            base = cpu_baseline(w, batch, samples=3, warmup=0, name_stage=varlen, target_s=1.0)
            base.pop("ms_per_step"), base.pop("packets_per_step")
            out["cpu_baseline"] = base
    del d_pk, d_v, d_offs, h_pk, h_v
    if world > 1:
        e.comm_destroy()
    e.close()
    torch.cuda.empty_cache()
    return out


def run_gpu(args) -> None:
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the CUDA library is the product, there is no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    if args.sweep:                                                 # BASELINE.json configs[4]: rule-table sweep 2^4..2^20, frac per size
        rows = []
        for k in SWEEP_SIZES:
            m = measure(args, f"c5:{k}", full=False)
            if rank == 0:
                rows.append({"entries": k, "value": m["value"], "frac": m["roofline"]["frac"], "kernel_ms": m["roofline"]["kernel_ms"],
                             "table_rows": m["config"]["table_rows"], "parity": m["parity"], "e2e": m["e2e"]["value"]})
        if rank == 0:
            print(json.dumps({"metric": METRIC, "workload": "c5 rule-table sweep (64 B flood, 100M packets/GPU)", "n_gpus": world, "unit": "Mpackets/s", "sweep": rows}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    main = measure(args, args.workload, full=True)
    also = {}
    for name in [x for x in args.also.split(",") if x and x != args.workload]:
        m = measure(args, name, full=False)
        if rank == 0:
            also[name] = {k: m[k] for k in ("value", "unit", "ms_per_step", "roofline", "e2e", "e2e_compact", "parity", "gpu_launches") if k in m}
            also[name]["workload"] = m["config"]["workload"]
    if rank == 0:
        line = {
            "metric": METRIC, "value": main["value"], "unit": "Mpackets/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        }
        line.update({k: v for k, v in main.items() if k not in ("value", "unit", "ms_per_step")})
        line["host_binding"] = numa
        if also:
            line["also"] = also
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", help="c2 | c2v6 | c3 | c3h | c4 | c5[:entries]   (default: c3, the north-star configuration)")
    ap.add_argument("--also", default=None, help="comma-separated workloads measured in the same run and reported under \"also\" (default: c2 next to c3)")
    ap.add_argument("--sweep", action="store_true", help="c5 rule-table sweep 2^4..2^20 entries: value and roofline fraction per size")
    ap.add_argument("--packets", type=int, default=0, help="override packets per GPU per step")
    ap.add_argument("--untiled", action="store_true", help="generate every packet of the batch instead of tiling `gen` unique ones (slow to generate; shows the tiling does not matter)")
    ap.add_argument("--seconds", type=float, default=3.0, help="length of the sustained leg (0 = skip)")
    ap.add_argument("--ref-sample-seconds", type=float, default=2.0, help="length of one CPU sample of the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--reserved-sms", type=int, default=-1, help="SMs the classify kernels leave to the side-stream collective (default: 4 when N > 1, else 0)")
    ap.add_argument("--metrics-capacity", type=int, default=131072, help="metrics_map max_entries per GPU (reference: 16384; the synthetic capture produces ~76k distinct keys)")
    args = ap.parse_args()
    if args.workload not in WORKLOADS and not args.workload.startswith("c5:"):
        ap.error(f"unknown workload {args.workload}")
    if args.also is None:
        args.also = "c2" if args.workload == "c3" else ""
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
