"""Host-side pieces of bench.py that can run without a GPU: the clock sampler's two sources and the reference arm."""
import json
import subprocess
import sys
import time
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def _fake_nvml(fail_after_init=False, reasons_mask=0x4):
    N = types.ModuleType("pynvml")
    N.NVML_CLOCK_SM = 1
    N.nvmlClocksEventReasonHwSlowdown, N.nvmlClocksEventReasonHwThermalSlowdown = 0x8, 0x40
    N.nvmlClocksEventReasonSwThermalSlowdown, N.nvmlClocksEventReasonSwPowerCap = 0x20, 0x4
    N.calls = 0
    N.nvmlInit = lambda: None
    N.nvmlDeviceGetHandleByIndex = lambda i: ("h", i)
    N.nvmlDeviceGetMaxClockInfo = lambda h, t: 1965

    def clock(h, t):
        N.calls += 1
        if fail_after_init and N.calls > 1:
            raise RuntimeError("nvml hiccup")
        return 1950 + (N.calls % 3) * 5
    N.nvmlDeviceGetClockInfo = clock
    N.nvmlDeviceGetCurrentClocksEventReasons = lambda h: reasons_mask
    return N


def test_clock_sampler_nvml_source(monkeypatch):
    monkeypatch.setitem(sys.modules, "pynvml", _fake_nvml())
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.02)
    assert s.rows == []                                      # parked until armed
    s.arm()
    time.sleep(0.05)
    c = s.stop()
    assert c["source"] == "nvml" and c["samples"] >= 5 and c["sm_max_mhz"] == 1965 and 1950 <= c["sm_mhz"] <= 1960
    assert c["reasons"] == ["sw_power_cap"]
    monkeypatch.setitem(sys.modules, "pynvml", _fake_nvml(fail_after_init=True, reasons_mask=0x8 | 0x20))   # the probe in __init__ is call 1
    s = bench.ClockSampler(0)
    s.start()
    s.arm()
    time.sleep(0.03)
    c = s.stop()                                             # later samples fail: the thread survives, nothing is recorded
    assert c["reasons"] == ["unavailable"]


def test_clock_sampler_falls_back_without_nvml(monkeypatch):
    broken = types.ModuleType("pynvml")

    def boom():
        raise RuntimeError("no NVML here")
    broken.nvmlInit = boom
    monkeypatch.setitem(sys.modules, "pynvml", broken)
    s = bench.ClockSampler(0)
    s.start()
    s.arm()
    time.sleep(0.05)
    c = s.stop()
    assert s.source == "nvidia-smi" and c["reasons"] == ["unavailable"]       # no GPU in this container: the query returns nothing


def test_reference_arm_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1", "--ref-sample-seconds", "0.2"],
                       capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "Mpackets/s" and d["value"] > 0
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["config"]["workload"] == bench.WORKLOADS["c3"]["desc"] and d["steps"] == 3 and d["warmup"] == 1        # the north-star configuration is the default
    assert d["cpu_baseline"]["samples"] == 3 and d["cpu_baseline"]["min"] <= d["value"] <= d["cpu_baseline"]["max"]
    assert set(d["config"]) == {"workload", "packets_per_gpu_per_step", "egress_rules", "table_rows", "layout", "synthetic", "l2_policy", "parallelism"}
    assert d["config"]["packets_per_gpu_per_step"] == 100_000_000 and d["config"]["table_rows"]["domains"] == 10000


def test_clock_sampler_never_armed_stops_cleanly():
    s = bench.ClockSampler(0)
    s.start()
    assert s.stop()["reasons"] == ["unavailable"] and not s.is_alive()


# ---- dry run of bench.run_gpu on CPU: CUDA, NCCL and the engine replaced by fakes, to catch host-side slips -----------
class _FakeStream:
    cuda_stream = 1

    def __init__(self, device=None):
        pass

    def wait_stream(self, other):
        pass


class _FakeEvent:
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass

    def elapsed_time(self, other):
        return 0.25


class _FakeStats:
    kernel_launches = 0
    metrics_dropped = 0


class _FakeEngine:
    def __init__(self, **kw):
        self.kw, self.st = kw, _FakeStats()

    def load_tables(self, t): pass
    def compile_rules(self, r): pass
    def publish(self): pass
    def counters_block_rows(self): return 8
    def totals_rows(self): return 4
    def map_count(self, i): return 1

    def stats(self):
        return self.st

    def _launch(self, *a):
        self.st.kernel_launches += 1
    classify_slots_dev = classify_varlen_dev = totals_export = counters_export = allgather_counters = _launch

    def comm_init(self, uid, rank, world): self.world = world
    def comm_destroy(self): pass
    def allgather_metrics(self, stream=0): pass

    def totals_global(self):
        return self.totals_merge(0, getattr(self, "world", 1))

    def counters_import(self, *a): pass
    def counters_reset(self): pass
    def close(self): pass

    def classify_slots(self, pk, stride, flags, out=None): pass
    def classify_hdr16(self, recs, aux, flags, out=None): pass
    def classify_varlen(self, pk, offs, flags, out=None): pass

    def totals_merge(self, ptr, world):
        import numpy as np
        from clawker_b200.abi import TOTAL_DTYPE
        t = np.zeros(1, dtype=TOTAL_DTYPE)
        t[0] = (7, (5, 3, 0))
        return t

    def counters(self, merged=False):
        import numpy as np
        from clawker_b200.abi import METRIC_DTYPE
        m = np.zeros(2, dtype=METRIC_DTYPE)
        m[0] = (7, 1, 443, 0, 0, 5)
        m[1] = (7, 2, 80, 1, 0, 3)
        return m


def _dry_run(monkeypatch, capsys, world, workload):
    import contextlib
    import numpy as np
    import torch
    import torch.distributed as dist
    from clawker_b200 import engine
    from clawker_b200 import abi as abi_mod
    cpu = torch.device("cpu")
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda i: None)
    monkeypatch.setattr(torch.cuda, "set_stream", lambda s: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "Stream", _FakeStream)
    monkeypatch.setattr(torch.cuda, "Event", _FakeEvent)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch, "device", lambda *a, **k: cpu)
    for name in ("init_process_group", "barrier", "destroy_process_group", "all_gather_into_tensor", "all_reduce", "broadcast_object_list"):
        monkeypatch.setattr(dist, name, lambda *a, **k: None)
    monkeypatch.setattr(engine, "comm_unique_id", lambda: b"\0" * 128)
    monkeypatch.setattr(engine, "pack_hdr16", lambda pk, stride=64, pinned=False: (np.zeros(pk.size // stride, dtype=abi_mod.HDR16_DTYPE), np.zeros(0, np.uint8)))
    monkeypatch.setattr(engine, "Engine", _FakeEngine)
    monkeypatch.setattr(engine, "pinned_empty", lambda n: np.empty(n, dtype=np.uint8))
    monkeypatch.setitem(bench.WORKLOADS[workload], "gen", 4096)
    monkeypatch.setenv("WORLD_SIZE", str(world))
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("LOCAL_RANK", "0")
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", str(world), "--steps", "4", "--warmup", "3", "--workload", workload,
                                      "--packets", "8192", "--no-cpu-baseline", "--no-parity", "--seconds", "0", "--also", ""])
    bench.main()
    return json.loads(capsys.readouterr().out.strip().splitlines()[-1])


def test_run_gpu_flow_single(monkeypatch, capsys):
    d = _dry_run(monkeypatch, capsys, 1, "c2")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["gpu_launches"] == 4 and d["config"]["parallelism"] == "single GPU"
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and d["clocks"]["reasons"] == ["unavailable"]
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["h2d_bytes_per_step"] == 8192 * 64
    assert d["e2e_compact"]["h2d_bytes_per_step"] == 8192 * 16 and d["e2e_compact"]["bytes_per_packet_over_pcie"] == 20.0
    # the arithmetic of the line, with the fake events' 0.25 ms between any two records: 4 timed steps take 0.25 ms in all
    assert d["ms_per_step"] == 0.0625 and d["value"] == round(8192 / 0.0625 / 1e3, 1)
    r = d["roofline"]
    assert r["algorithmic_bytes_per_packet"] == 68.0 and r["kernel_ms"] == 0.25                   # one classify call per step, 0.25 ms each
    assert abs(r["achieved"] - 8192 * 68 / 0.25e-3 / 1e9) < 0.06 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3 and r["bound"] == "hbm" and r["unit"] == "GB/s"


def test_run_gpu_flow_multi_and_varlen(monkeypatch, capsys):
    d = _dry_run(monkeypatch, capsys, 2, "c2")
    assert d["n_gpus"] == 2 and d["gpu_launches"] == 8 and "4 SMs reserved" in d["config"]["collective"]     # classify + totals per step
    d = _dry_run(monkeypatch, capsys, 1, "c3h")
    assert "k_varlen_head" in d["roofline"]["kernel"] and "e2e_compact" not in d and d["config"]["layout"].startswith("varlen")


# ---- bench.check_parity on CPU: does the checker itself catch a wrong word, a wrong copy, a wrong counter? -------------
class _ParityEngine:
    def __init__(self, rows, dropped=0):
        self.rows, self.dropped = rows, dropped

    def counters_reset(self): pass
    def counters(self): return self.rows

    def stats(self):
        s = _FakeStats()
        s.metrics_dropped = self.dropped
        return s


def test_check_parity_catches_words_copies_and_counters(monkeypatch):
    import numpy as np
    import torch
    from clawker_b200 import capture
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    w = capture.make_workload("c2")
    gen, reps = 20_000, 3
    pk = capture.gen_slots(w, gen, seed=9)
    o, _ = bench._oracle_for(w, False, want_cidr=True)
    want, wm = o.classify(pk, None, flags=0)
    rows = wm.copy()
    rows["count"] *= reps                                     # what a correct engine holds after one pass over the tiled batch
    good = torch.from_numpy(np.tile(want, reps).view(np.int32).copy())
    args = dict(w=w, batch=(pk, None), n=gen * reps, gen=gen, varlen=False, flags=0, run_step=lambda: None, threads=2)
    p = bench.check_parity(_ParityEngine(rows), d_v=good, **args)
    assert p["mismatches"] == 0 and p["copy_mismatches"] == 0 and p["copies_checked"] == reps and p["rate"] == 1.0
    assert p["counters"] == {"rows": int(wm.size), "oracle_rows": int(wm.size), "mismatches": 0}
    bad = good.clone(); bad[5] ^= 1                           # a wrong word in the first copy: the oracle comparison sees it (and every other copy now differs from it)
    p = bench.check_parity(_ParityEngine(rows), d_v=bad, **args)
    assert p["mismatches"] == 1 and p["copy_mismatches"] == reps - 1 and p["rate"] < 1.0
    bad = good.clone(); bad[gen + 7] ^= 4                     # a wrong word in the MIDDLE copy: only the device-side copy comparison can see it
    p = bench.check_parity(_ParityEngine(rows), d_v=bad, **args)
    assert p["mismatches"] == 0 and p["copy_mismatches"] == 1 and p["rate"] < 1.0
    r2 = rows.copy(); r2["count"][3] += 1                     # one increment too many
    assert bench.check_parity(_ParityEngine(r2), d_v=good, **args)["counters"]["mismatches"] == 1
    r3 = np.concatenate([rows, rows[:1]]); r3["dst_port"][-1] ^= 0x5555        # a stray row (what a header read after a TMA refill produced in round 2)
    assert bench.check_parity(_ParityEngine(r3), d_v=good, **args)["counters"]["mismatches"] >= 1
    assert "skipped" in bench.check_parity(_ParityEngine(rows, dropped=5), d_v=good, **args)["counters"]
