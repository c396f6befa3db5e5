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
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "Mpackets/s" and d["value"] > 0
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["config"]["workload"] == bench.WORKLOADS["c2"]["desc"] and d["steps"] == 2 and d["warmup"] == 1


def test_clock_sampler_never_armed_stops_cleanly():
    s = bench.ClockSampler(0)
    s.start()
    assert s.stop()["reasons"] == ["unavailable"] and not s.is_alive()
