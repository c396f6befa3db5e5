"""ctypes driver for the CPU oracle (TEST INFRASTRUCTURE — see oracle/README.md).

`Oracle("port")` loads oracle/liboracle_port.so (the C restatement);
`Oracle("ref")` loads oracle/_ref/liboracle_ref.so (the reference's own clawker.c on the host).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent

ORC_F_CIDR_LINEAR = 0x100


class ContainerConfig(C.Structure):
    _fields_ = [("envoy_ip", C.c_uint32), ("coredns_ip", C.c_uint32), ("gateway_ip", C.c_uint32),
                ("net_addr", C.c_uint32), ("net_mask", C.c_uint32), ("host_proxy_ip", C.c_uint32),
                ("host_proxy_port", C.c_uint16), ("egress_port", C.c_uint16)]


class Endpoint(C.Structure):
    _fields_ = [("cgroup_id", C.c_uint64), ("uid", C.c_uint32), ("so_mark", C.c_uint32)]


EVENT_DTYPE = np.dtype([("cgroup_id", "<u8"), ("uid", "<u4"), ("so_mark", "<u4"), ("user_ip4", "<u4"),
                        ("user_ip6", "<u4", (4,)), ("user_port", "<u2"), ("hook", "u1"), ("sock_type", "u1")])
METRIC_DTYPE = np.dtype([("cgroup_id", "<u8"), ("domain_hash", "<u4"), ("dst_port", "<u2"), ("action", "u1"),
                         ("_pad", "u1"), ("count", "<u8")])
DNSUPD_DTYPE = np.dtype([("ip", "<u4"), ("domain_hash", "<u4"), ("ttl", "<u4"), ("packet", "<u4")])
assert EVENT_DTYPE.itemsize == 40 and METRIC_DTYPE.itemsize == 24


def build(force: bool = False) -> None:
    """Compile the oracle libraries (gcc).  The `_ref` library is only (re)built where
    /root/reference exists; on the GPU box the prebuilt file that travelled is used."""
    port = HERE / "liboracle_port.so"
    need = force or not port.exists() or port.stat().st_mtime < (HERE / "clk_oracle.c").stat().st_mtime
    ref_src = Path("/root/reference/internal/ebpf/bpf/clawker.c")
    if ref_src.exists() and not (HERE / "_ref" / "liboracle_ref.so").exists():
        need = True
    if need:
        subprocess.run(["make", "-C", str(HERE)], check=True, capture_output=True)


def have_ref() -> bool:
    return (HERE / "_ref" / "liboracle_ref.so").exists()


class Oracle:
    def __init__(self, kind: str = "port"):
        build()
        path = HERE / ("liboracle_port.so" if kind == "port" else "_ref/liboracle_ref.so")
        if not path.exists():
            raise FileNotFoundError(path)
        L = self.lib = C.CDLL(str(path))
        self.kind = kind
        L.orc_tables_new.restype = C.c_void_p
        L.orc_metrics_new.restype = C.c_void_p
        L.orc_kind.restype = C.c_char_p
        L.orc_domain_hash.restype = C.c_uint32
        L.orc_domain_hash.argtypes = [C.c_char_p, C.c_uint32]
        L.orc_tables_free.argtypes = [C.c_void_p]
        L.orc_metrics_free.argtypes = [C.c_void_p]
        L.orc_set_container.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(ContainerConfig)]
        L.orc_set_bypass.argtypes = [C.c_void_p, C.c_uint64, C.c_uint8]
        L.orc_set_dns.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_set_route.argtypes = [C.c_void_p, C.c_uint32, C.c_uint16, C.c_uint16]
        L.orc_set_endpoint.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(Endpoint)]
        L.orc_set_endpoint6.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(Endpoint)]
        L.orc_add_cidr.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_add_zone.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_add_sni.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_add_http_vhost.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.orc_classify_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64,
                                         C.c_uint32, C.c_void_p, C.c_int]
        L.orc_classify_events.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]
        L.orc_classify_batch_timed.restype = C.c_double
        L.orc_classify_batch_timed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32, C.c_void_p, C.c_int]
        L.orc_classify_batch_timed2.restype = C.c_double
        L.orc_classify_batch_timed2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_metrics_count.restype = C.c_uint32
        L.orc_metrics_count.argtypes = [C.c_void_p]
        L.orc_metrics_dump.restype = C.c_uint32
        L.orc_metrics_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_dns_answers.restype = C.c_uint64
        L.orc_dns_answers.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        assert L.orc_kind().decode() == ("port" if kind == "port" else "reference")
        self.t = L.orc_tables_new()

    def close(self):
        if self.t:
            self.lib.orc_tables_free(self.t)
            self.t = None

    # ---- tables -----------------------------------------------------------------------
    def domain_hash(self, s: str) -> int:
        b = s.encode()
        return self.lib.orc_domain_hash(b, len(b))

    def load(self, tables: "dict") -> "Oracle":
        """Load a table set in the neutral dict form produced by clawker_b200.capture.make_tables()."""
        L = self.lib
        for cg, cfg in tables.get("containers", {}).items():
            c = ContainerConfig(*cfg)
            L.orc_set_container(self.t, cg, C.byref(c))
        for cg, v in tables.get("bypass", {}).items():
            L.orc_set_bypass(self.t, cg, v)
        for ip, (h, exp) in tables.get("dns_cache", {}).items():
            L.orc_set_dns(self.t, ip, h, exp)
        for (h, port), eport in tables.get("routes", {}).items():
            L.orc_set_route(self.t, h, port, eport)
        for ip, (cg, uid, mark) in tables.get("endpoints", {}).items():
            e = Endpoint(cg, uid, mark)
            L.orc_set_endpoint(self.t, ip, C.byref(e))
        for ip6, (cg, uid, mark) in tables.get("endpoints6", {}).items():
            e = Endpoint(cg, uid, mark)
            L.orc_set_endpoint6(self.t, bytes(ip6), C.byref(e))
        for addr, mask in tables.get("cidrs", []):
            L.orc_add_cidr(self.t, addr, mask)
        for z in tables.get("zones", []):
            L.orc_add_zone(self.t, z.encode())
        for s in tables.get("snis", []):
            L.orc_add_sni(self.t, s.encode())
        for domains, routes in tables.get("http", []):                 # [(domains, [(prefix, allow), ...]), ...] in config order
            L.orc_add_http_vhost(self.t, "\n".join(domains).encode(), "\n".join(("A" if a else "D") + p for p, a in routes).encode())
        return self

    # ---- classification ---------------------------------------------------------------
    def _metrics_out(self, m):
        n = self.lib.orc_metrics_count(m)
        out = np.zeros(n, dtype=METRIC_DTYPE)
        if n:
            self.lib.orc_metrics_dump(m, out.ctypes.data, n)
        self.lib.orc_metrics_free(m)
        return out

    def classify(self, pkts: np.ndarray, offs16: np.ndarray | None = None, stride: int = 64, flags: int = 0,
                 nthreads: int = 1, metrics: bool = True):
        pkts = np.ascontiguousarray(pkts, dtype=np.uint8)
        if offs16 is None:
            n = pkts.size // stride
            op = None
        else:
            offs16 = np.ascontiguousarray(offs16, dtype=np.uint32)
            n = offs16.size - 1
            op = offs16.ctypes.data
        verd = np.zeros(n, dtype=np.uint32)
        m = self.lib.orc_metrics_new() if metrics else None
        self.lib.orc_classify_batch(self.t, m, pkts.ctypes.data, op, stride, n, flags, verd.ctypes.data, nthreads)
        return verd, (self._metrics_out(m) if metrics else None)

    def classify_timed(self, pkts: np.ndarray, offs16: np.ndarray | None = None, stride: int = 64, flags: int = 0,
                       nthreads: int = 1, verd: np.ndarray | None = None) -> float:
        """CPU-baseline form: seconds spent in the parallel section only (threads and their private PERCPU metrics
        maps exist before the clock starts; the per-thread maps are discarded, not merged).  `verd` receives the words."""
        pkts = np.ascontiguousarray(pkts, dtype=np.uint8)
        if offs16 is None:
            n, op = pkts.size // stride, None
        else:
            offs16 = np.ascontiguousarray(offs16, dtype=np.uint32)
            n, op = offs16.size - 1, offs16.ctypes.data
        if verd is None:
            verd = np.zeros(n, dtype=np.uint32)
        assert verd.size >= n and verd.dtype == np.uint32
        return float(self.lib.orc_classify_batch_timed(self.t, pkts.ctypes.data, op, stride, n, flags, verd.ctypes.data, nthreads))

    def classify_timed2(self, pkts: np.ndarray, offs16: np.ndarray | None = None, stride: int = 64, flags: int = 0,
                        nthreads: int = 1, repeat: int = 1, pin: bool = True, verd: np.ndarray | None = None) -> float:
        """Stabilised CPU-baseline form: threads pinned one per allowed CPU, every thread classifies a private,
        self-allocated copy of its slice `repeat` times inside the timed section.  Seconds of the parallel section."""
        pkts = np.ascontiguousarray(pkts, dtype=np.uint8)
        if offs16 is None:
            n, op = pkts.size // stride, None
        else:
            offs16 = np.ascontiguousarray(offs16, dtype=np.uint32)
            n, op = offs16.size - 1, offs16.ctypes.data
        if verd is None:
            verd = np.zeros(n, dtype=np.uint32)
        assert verd.size >= n and verd.dtype == np.uint32
        return float(self.lib.orc_classify_batch_timed2(self.t, pkts.ctypes.data, op, stride, n, flags, verd.ctypes.data,
                                                        nthreads, repeat, 1 if pin else 0))

    def classify_events(self, events: np.ndarray, flags: int = 0):
        events = np.ascontiguousarray(events, dtype=EVENT_DTYPE)
        verd = np.zeros(events.size, dtype=np.uint32)
        m = self.lib.orc_metrics_new()
        self.lib.orc_classify_events(self.t, m, events.ctypes.data, events.size, flags, verd.ctypes.data)
        return verd, self._metrics_out(m)

    def dns_answers(self, pkts: np.ndarray, offs16: np.ndarray, cap: int | None = None):
        pkts = np.ascontiguousarray(pkts, dtype=np.uint8)
        offs16 = np.ascontiguousarray(offs16, dtype=np.uint32)
        n = offs16.size - 1
        cap = cap if cap is not None else n * 31
        out = np.zeros(cap, dtype=DNSUPD_DTYPE)
        k = self.lib.orc_dns_answers(self.t, pkts.ctypes.data, offs16.ctypes.data, n, out.ctypes.data, cap)
        return out[:k]


def sort_metrics(m: np.ndarray) -> np.ndarray:
    """Canonical order for comparing two metric tables."""
    if m.size == 0:
        return m
    return np.sort(m, order=["cgroup_id", "domain_hash", "dst_port", "action"])
