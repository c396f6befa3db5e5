"""ctypes binding of libclawker_gpu.so (include/clawker_gpu.h).

The library is the product; this module only marshals numpy buffers into its C-ABI so the Python
mirrors of the reference's Go packages (clawker_b200.ebpf / .dnsbpf / .firewall), the tests and
bench.py can drive it.  There is no fallback: a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from . import abi
from .abi import (Cfg, ContainerConfig, DNSEntry, Endpoint, Cidr, Ports, Route, RouteKey, RouteVal, Rule, Stats,
                  DNSUPD_DTYPE, EVENT_DTYPE, METRIC_DTYPE)

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["CLK_LIB"]) if os.environ.get("CLK_LIB") else PKG / "libclawker_gpu.so"
CLI_PATH = PKG / "ebpf-manager"                   # control-plane verbs on the pinned state (csrc/ebpf_manager.cpp)


class ClkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"clk error {code}: {msg}")
        self.code = code


def build_library(force: bool = False) -> Path:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> clawker_b200/libclawker_gpu.so (in-tree)."""
    srcs = list((PKG / "csrc").glob("*")) + [PKG.parent / "include" / "clawker_gpu.h"]
    stale = not LIB_PATH.exists() or not CLI_PATH.exists() or any(s.stat().st_mtime > LIB_PATH.stat().st_mtime for s in srcs)
    if force or stale:
        r = subprocess.run(["make", "-C", str(PKG / "csrc")] + (["-B"] if force else []), capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building libclawker_gpu.so failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise FileNotFoundError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(the CUDA library is the product; there is no CPU fallback)")
    L = C.CDLL(str(LIB_PATH))
    vp, u32, u64, i32, sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_size_t
    sig = {
        "clk_abi_version": (u32, []),
        "clk_strerror": (C.c_char_p, [i32]),
        "clk_last_error": (C.c_char_p, [vp]),
        "clk_engine_create": (i32, [C.POINTER(Cfg), C.POINTER(vp)]),
        "clk_engine_destroy": (None, [vp]),
        "clk_map_update": (i32, [vp, i32, vp, sz, vp, sz]),
        "clk_map_delete": (i32, [vp, i32, vp, sz]),
        "clk_map_lookup": (i32, [vp, i32, vp, sz, vp, sz]),
        "clk_map_iter": (i32, [vp, i32, vp, vp, u32, C.POINTER(u32)]),
        "clk_map_count": (i32, [vp, i32, C.POINTER(u32)]),
        "clk_routes_replace": (i32, [vp, vp, u32]),
        "clk_dns_update": (i32, [vp, u32, u32, u32, u32]),
        "clk_dns_gc": (i32, [vp, u32]),
        "clk_validate_dst": (i32, [C.c_char_p]),
        "clk_domain_hash": (u32, [C.c_char_p, sz]),
        "clk_ip_to_u32": (i32, [C.c_char_p, C.POINTER(u32)]),
        "clk_cidr_to_addr_mask": (i32, [C.c_char_p, C.POINTER(u32), C.POINTER(u32)]),
        "clk_rules_compile": (i32, [vp, vp, u32, C.POINTER(Ports), C.POINTER(u32)]),
        "clk_rules_dump": (i32, [vp, i32, C.c_char_p, sz]),
        "clk_publish": (i32, [vp]),
        "clk_classify_slots": (i32, [vp, vp, u32, u64, u32, vp]),
        "clk_classify_varlen": (i32, [vp, vp, vp, u64, u32, vp]),
        "clk_classify_slots_dev": (i32, [vp, vp, u32, u64, u32, vp, vp]),
        "clk_classify_varlen_dev": (i32, [vp, vp, vp, u64, u32, vp, vp]),
        "clk_classify_events": (i32, [vp, vp, u64, u32, vp]),
        "clk_pack_hdr16": (i32, [vp, u32, u64, vp, vp, u64, C.POINTER(u64)]),
        "clk_classify_hdr16": (i32, [vp, vp, u64, vp, u64, u32, vp]),
        "clk_classify_hdr16_dev": (i32, [vp, vp, u64, vp, u32, vp, vp]),
        "clk_dns_answers": (i32, [vp, vp, vp, u64, vp, vp, u64, C.POINTER(u64), i32, u32]),
        "clk_counters_read": (i32, [vp, vp, u32, C.POINTER(u32)]),
        "clk_counters_read_global": (i32, [vp, vp, u32, C.POINTER(u32)]),
        "clk_counters_reset": (i32, [vp]),
        "clk_counters_block_rows": (i32, [vp, C.POINTER(u32)]),
        "clk_counters_export": (i32, [vp, vp, vp]),
        "clk_counters_import": (i32, [vp, vp, u32, vp]),
        "clk_pcap_pack": (i32, [C.c_char_p, C.POINTER(u64), vp, u64, vp, u64, C.POINTER(u64)]),
        "clk_classify_pcap": (i32, [vp, C.c_char_p, u32, vp, u64, C.POINTER(u64)]),
        "clk_pcap_last_error": (C.c_char_p, []),
        "clk_afpacket_open": (i32, [C.c_char_p, u32, C.POINTER(vp)]),
        "clk_afpacket_read": (i32, [vp, vp, u64, vp, u64, i32, C.POINTER(u64)]),
        "clk_afpacket_read_hdr16": (i32, [vp, vp, u64, vp, u64, i32, C.POINTER(u64), C.POINTER(u64)]),
        "clk_afpacket_close": (i32, [vp]),
        "clk_tcp_coalesce": (i32, [vp, vp, u64, vp, u64, vp, C.POINTER(u64)]),
        "clk_ingest_last_error": (C.c_char_p, []),
        "clk_state_save": (i32, [vp, C.c_char_p]),
        "clk_state_load": (i32, [vp, C.c_char_p]),
        "clk_totals_rows": (i32, [vp, C.POINTER(u32)]),
        "clk_totals_export": (i32, [vp, vp, vp]),
        "clk_totals_merge": (i32, [vp, vp, u32, vp, u32, C.POINTER(u32)]),
        "clk_comm_unique_id": (i32, [vp, sz, C.POINTER(sz)]),
        "clk_comm_init": (i32, [vp, vp, sz, i32, i32]),
        "clk_comm_destroy": (i32, [vp]),
        "clk_comm_info": (i32, [vp, C.POINTER(i32), C.POINTER(i32)]),
        "clk_allgather_counters": (i32, [vp, vp]),
        "clk_totals_read_global": (i32, [vp, vp, u32, C.POINTER(u32)]),
        "clk_allgather_metrics": (i32, [vp, vp]),
        "clk_get_stats": (i32, [vp, C.POINTER(Stats)]),
        "clk_sync": (i32, [vp]),
        "clk_host_alloc": (i32, [sz, C.POINTER(vp)]),
        "clk_host_free": (i32, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


EXPORTS = ["clk_abi_version", "clk_strerror", "clk_last_error", "clk_engine_create", "clk_engine_destroy",
           "clk_map_update", "clk_map_delete", "clk_map_lookup", "clk_map_iter", "clk_map_count", "clk_routes_replace",
           "clk_dns_update", "clk_dns_gc", "clk_validate_dst", "clk_domain_hash", "clk_ip_to_u32",
           "clk_cidr_to_addr_mask", "clk_rules_compile", "clk_rules_dump", "clk_publish", "clk_classify_slots",
           "clk_classify_varlen", "clk_classify_slots_dev", "clk_classify_varlen_dev", "clk_classify_events", "clk_pack_hdr16", "clk_classify_hdr16", "clk_classify_hdr16_dev",
           "clk_dns_answers", "clk_counters_read", "clk_counters_read_global", "clk_counters_reset", "clk_counters_block_rows",
           "clk_counters_export", "clk_counters_import", "clk_totals_rows", "clk_totals_export", "clk_totals_merge", "clk_comm_unique_id", "clk_comm_init", "clk_comm_destroy", "clk_comm_info",
           "clk_allgather_counters", "clk_totals_read_global", "clk_allgather_metrics", "clk_state_save", "clk_state_load", "clk_pcap_pack", "clk_classify_pcap", "clk_pcap_last_error", "clk_afpacket_open", "clk_afpacket_read", "clk_afpacket_read_hdr16", "clk_afpacket_close", "clk_tcp_coalesce", "clk_ingest_last_error", "clk_get_stats", "clk_sync", "clk_host_alloc", "clk_host_free"]


class PacketRing:
    """AF_PACKET / TPACKET_V3 receive ring on one interface (clk_afpacket_*): read() returns (blob, offs16) in the varlen layout."""

    def __init__(self, ifname: str, ring_mib: int = 0):
        self.L, self.h = lib(), C.c_void_p()
        rc = self.L.clk_afpacket_open(ifname.encode(), ring_mib, C.byref(self.h))
        if rc:
            raise ClkError(rc, self.L.clk_ingest_last_error().decode())

    def read(self, cap_bytes: int = 16 << 20, cap_frames: int = 1 << 16, timeout_ms: int = 100):
        blob, offs, n = np.zeros(cap_bytes, dtype=np.uint8), np.zeros(cap_frames + 1, dtype=np.uint32), C.c_uint64()
        rc = self.L.clk_afpacket_read(self.h, blob.ctypes.data, cap_bytes, offs.ctypes.data, cap_frames, timeout_ms, C.byref(n))
        if rc:
            raise ClkError(rc, self.L.clk_ingest_last_error().decode())
        return blob[: int(offs[n.value]) * 16], offs[: n.value + 1]

    def read_hdr16(self, cap_frames: int = 1 << 16, aux_cap: int = 4096, timeout_ms: int = 100):
        """(recs, aux) for Engine.classify_hdr16: compact ingest records straight from the ring."""
        from .abi import HDR16_DTYPE
        recs, aux = np.zeros(cap_frames, dtype=HDR16_DTYPE), np.zeros(aux_cap * 64, dtype=np.uint8)
        n, na = C.c_uint64(), C.c_uint64()
        rc = self.L.clk_afpacket_read_hdr16(self.h, recs.ctypes.data, cap_frames, aux.ctypes.data, aux_cap, timeout_ms, C.byref(n), C.byref(na))
        if rc:
            raise ClkError(rc, self.L.clk_ingest_last_error().decode())
        return recs[: n.value], aux[: na.value * 64]

    def close(self):
        if self.h:
            self.L.clk_afpacket_close(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def tcp_coalesce(blob: np.ndarray, offs16: np.ndarray):
    """clk_tcp_coalesce: (blob, offs16, n_merged) with split ClientHellos / HTTP request heads folded into their first segment."""
    L = lib()
    blob, offs16 = np.ascontiguousarray(blob, dtype=np.uint8), np.ascontiguousarray(offs16, dtype=np.uint32)
    n = offs16.size - 1
    out, oo, merged = np.zeros(2 * blob.size + 64, dtype=np.uint8), np.zeros(n + 1, dtype=np.uint32), C.c_uint64()
    rc = L.clk_tcp_coalesce(blob.ctypes.data, offs16.ctypes.data, n, out.ctypes.data, out.size, oo.ctypes.data, C.byref(merged))
    if rc:
        raise ClkError(rc, L.clk_ingest_last_error().decode())
    return out[: int(oo[n]) * 16], oo, merged.value


def pcap_pack(path: str, cap_bytes: int = 64 << 20, cap_frames: int = 1 << 20):
    """Generator over a classic pcap file: (blob, offs16) chunks in the varlen layout (clk_pcap_pack)."""
    L = lib()
    cursor = C.c_uint64(0)
    while True:
        blob, offs, n = np.zeros(cap_bytes, dtype=np.uint8), np.zeros(cap_frames + 1, dtype=np.uint32), C.c_uint64()
        rc = L.clk_pcap_pack(os.fsencode(path), C.byref(cursor), blob.ctypes.data, cap_bytes, offs.ctypes.data, cap_frames, C.byref(n))
        if rc:
            raise ClkError(rc, L.clk_pcap_last_error().decode())
        if n.value:
            yield blob[: int(offs[n.value]) * 16], offs[: n.value + 1]
        elif cursor.value:
            raise ClkError(-7, "a frame does not fit the chunk buffers")
        if cursor.value == 0:
            return


def pinned_empty(nbytes: int) -> np.ndarray:
    """uint8 array backed by clk_host_alloc (page-locked) memory; kept alive by the array's base."""
    p = C.c_void_p()
    rc = lib().clk_host_alloc(nbytes, C.byref(p))
    if rc:
        raise ClkError(rc, lib().clk_last_error(None).decode())
    buf = (C.c_uint8 * nbytes).from_address(p.value)
    arr = np.frombuffer(buf, dtype=np.uint8)

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr

        def __del__(self):
            try:
                lib().clk_host_free(self.ptr)
            except Exception:
                pass
    _owners[arr.ctypes.data] = _Owner(p.value)
    return arr


_owners: dict = {}


class Engine:
    """One GPU's classifier instance (clk_engine)."""

    def __init__(self, device: int = -1, host_only: bool = False, max_containers: int = 0, max_dns_cache: int = 0,
                 max_routes: int = 0, max_metrics: int = 0, max_cidrs: int = 0, reserved_sms: int = 0):
        self.L = lib()
        cfg = Cfg(self.L.clk_abi_version(), device, max_containers, max_dns_cache, max_routes, max_metrics, max_cidrs,
                  1 if host_only else 0, reserved_sms)
        h = C.c_void_p()
        rc = self.L.clk_engine_create(C.byref(cfg), C.byref(h))
        if rc:
            raise ClkError(rc, self.L.clk_last_error(None).decode())
        self.h = h

    # ---- plumbing ---------------------------------------------------------------------------------
    def _ck(self, rc: int, ok=(0,)):
        if rc not in ok:
            raise ClkError(rc, self.L.clk_last_error(self.h).decode())
        return rc

    def close(self):
        if getattr(self, "h", None):
            self.L.clk_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- map ops ------------------------------------------------------------------------------------
    def map_update(self, map_id: int, key, val=None):
        vp = C.byref(val) if val is not None else None
        return self._ck(self.L.clk_map_update(self.h, map_id, C.byref(key), C.sizeof(key), vp, C.sizeof(val) if val is not None else 0))

    def map_delete(self, map_id: int, key) -> int:
        """Returns 0 or abi.ENOENT (ebpf.ErrKeyNotExist)."""
        return self._ck(self.L.clk_map_delete(self.h, map_id, C.byref(key), C.sizeof(key)), ok=(0, abi.ENOENT))

    def map_lookup(self, map_id: int, key, val):
        rc = self._ck(self.L.clk_map_lookup(self.h, map_id, C.byref(key), C.sizeof(key), C.byref(val), C.sizeof(val)), ok=(0, abi.ENOENT))
        return None if rc else val

    def map_count(self, map_id: int) -> int:
        n = C.c_uint32()
        self._ck(self.L.clk_map_count(self.h, map_id, C.byref(n)))
        return n.value

    _MAP_SIZES = {0: (8, 28), 1: (8, 1), 2: (4, 8), 3: (8, 4), 5: (4, 16), 6: (16, 16), 7: (8, 1)}

    def map_rows(self, map_id: int) -> dict:
        """Every row of a map as {key bytes: value bytes} (clk_map_iter)."""
        klen, vlen = self._MAP_SIZES[map_id]
        n = self.map_count(map_id)
        keys, vals, got = C.create_string_buffer(max(1, n * klen)), C.create_string_buffer(max(1, n * vlen)), C.c_uint32()
        self._ck(self.L.clk_map_iter(self.h, map_id, keys, vals, n, C.byref(got)))
        return {keys.raw[i * klen:(i + 1) * klen]: vals.raw[i * vlen:(i + 1) * vlen] for i in range(min(n, got.value))}

    # capture files
    def classify_pcap(self, path: str, flags: int = 0) -> np.ndarray:
        """Verdict word per frame of a classic pcap file (read, packed and classified in overlapped chunks)."""
        n = C.c_uint64()
        rc = self.L.clk_classify_pcap(self.h, os.fsencode(path), flags, None, 0, C.byref(n))      # count first
        if rc:
            raise ClkError(rc, self.L.clk_pcap_last_error().decode())
        out = np.empty(n.value, dtype=np.uint32)
        if n.value:
            rc = self.L.clk_classify_pcap(self.h, os.fsencode(path), flags, out.ctypes.data, n.value, C.byref(n))
            if rc:
                raise ClkError(rc, self.L.clk_pcap_last_error().decode())
        return out

    # the pinned state shared with the `ebpf-manager` verbs
    def state_save(self, path: str):
        self._ck(self.L.clk_state_save(self.h, os.fsencode(path)))

    def state_load(self, path: str):
        self._ck(self.L.clk_state_load(self.h, os.fsencode(path)))

    def routes_replace(self, routes):
        arr = (Route * len(routes))(*[Route(*r) for r in routes])
        self._ck(self.L.clk_routes_replace(self.h, arr, len(routes)))

    def dns_update(self, ip: int, domain_hash: int, ttl: int, now: int = 0):
        self._ck(self.L.clk_dns_update(self.h, ip, domain_hash, ttl, now))

    def dns_gc(self, now: int = 0) -> int:
        rc = self.L.clk_dns_gc(self.h, now)
        if rc < 0:
            self._ck(rc)
        return rc

    def load_tables(self, t: dict) -> "Engine":
        """Load a neutral table set (clawker_b200.capture.base_tables form)."""
        for cg, cfg in t.get("containers", {}).items():
            self.map_update(abi.MAP_CONTAINER, C.c_uint64(cg), ContainerConfig(*cfg))
        for cg, v in t.get("bypass", {}).items():
            self.map_update(abi.MAP_BYPASS, C.c_uint64(cg), C.c_uint8(v))
        for ip, (h, exp) in t.get("dns_cache", {}).items():
            self.map_update(abi.MAP_DNS_CACHE, C.c_uint32(ip), DNSEntry(h, exp))
        for (h, port), eport in t.get("routes", {}).items():
            self.map_update(abi.MAP_ROUTE, RouteKey(h, port, 0), RouteVal(eport, 0))
        for ip, (cg, uid, mark) in t.get("endpoints", {}).items():
            self.map_update(abi.MAP_ENDPOINT, C.c_uint32(ip), Endpoint(cg, uid, mark))
        for ip6, (cg, uid, mark) in t.get("endpoints6", {}).items():
            self.map_update(abi.MAP_ENDPOINT6, (C.c_uint8 * 16)(*ip6), Endpoint(cg, uid, mark))
        for addr, mask in t.get("cidrs", []):
            self.map_update(abi.MAP_CIDR, Cidr(addr, mask), None)
        return self

    # ---- rules --------------------------------------------------------------------------------------
    def compile_rules(self, rules, egress_port: int = 10000, tcp_port_base: int = 10001) -> int:
        """rules: iterable of dicts / objects with dst, proto, port, action.  Returns the warning count."""
        def g(r, k, d):
            return (r.get(k, d) if isinstance(r, dict) else getattr(r, k, d))
        arr = (Rule * len(rules))()
        keep = []
        for i, r in enumerate(rules):
            b = [str(g(r, "dst", "")).encode(), str(g(r, "proto", "") or "").encode(), str(g(r, "action", "") or "").encode(),
                 str(g(r, "path_default", "") or "").encode()]
            prs = [((x["path"], x.get("action", "")) if isinstance(x, dict) else (x[0], x[1])) for x in (g(r, "path_rules", None) or [])]
            pb = [(str(p).encode(), str(a or "").encode()) for p, a in prs]
            parr = (abi.PathRule * len(pb))(*[abi.PathRule(p, a) for p, a in pb]) if pb else None
            keep.append((b, pb, parr))
            arr[i] = Rule(b[0], b[1], int(g(r, "port", 0) or 0), b[2], b[3], parr, len(pb))
        w = C.c_uint32()
        ports = Ports(egress_port, tcp_port_base)
        self._ck(self.L.clk_rules_compile(self.h, arr, len(rules), C.byref(ports), C.byref(w)))
        return w.value

    def rules_dump(self, kind: int) -> list:
        n = self.L.clk_rules_dump(self.h, kind, None, 0)
        buf = C.create_string_buffer(n)
        self.L.clk_rules_dump(self.h, kind, buf, n)
        s = buf.value.decode()
        return [x for x in s.split("\n") if x]

    def publish(self):
        self._ck(self.L.clk_publish(self.h))

    # ---- classification -----------------------------------------------------------------------------
    def classify_slots(self, pkts: np.ndarray, stride: int = 64, flags: int = 0, out: np.ndarray | None = None) -> np.ndarray:
        pkts = np.ascontiguousarray(pkts, dtype=np.uint8)
        n = pkts.size // stride
        verd = out if out is not None else np.empty(n, dtype=np.uint32)
        self._ck(self.L.clk_classify_slots(self.h, pkts.ctypes.data, stride, n, flags, verd.ctypes.data))
        return verd

    def classify_varlen(self, blob: np.ndarray, offs16: np.ndarray, flags: int = 0, out: np.ndarray | None = None) -> np.ndarray:
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        offs16 = np.ascontiguousarray(offs16, dtype=np.uint32)
        n = offs16.size - 1
        verd = out if out is not None else np.empty(n, dtype=np.uint32)
        self._ck(self.L.clk_classify_varlen(self.h, blob.ctypes.data, offs16.ctypes.data, n, flags, verd.ctypes.data))
        return verd

    def classify_slots_dev(self, d_pkts: int, stride: int, n: int, flags: int, d_verdicts: int, stream: int = 0):
        self._ck(self.L.clk_classify_slots_dev(self.h, d_pkts, stride, n, flags, d_verdicts, stream or None))

    def classify_varlen_dev(self, d_pkts: int, d_offs16: int, n: int, flags: int, d_verdicts: int, stream: int = 0):
        self._ck(self.L.clk_classify_varlen_dev(self.h, d_pkts, d_offs16, n, flags, d_verdicts, stream or None))

    def classify_hdr16(self, recs: np.ndarray, aux: np.ndarray | None = None, flags: int = 0, out: np.ndarray | None = None) -> np.ndarray:
        """Compact ingest records (abi.HDR16_DTYPE, 16 B per frame) + the whole frames they point at (aux, 64 B slots)."""
        n = recs.shape[0]
        n_aux = 0 if aux is None else aux.size // 64
        verd = out if out is not None else np.empty(n, dtype=np.uint32)
        self._ck(self.L.clk_classify_hdr16(self.h, recs.ctypes.data, n, aux.ctypes.data if n_aux else None, n_aux, flags, verd.ctypes.data))
        return verd

    def classify_hdr16_dev(self, d_recs: int, n: int, d_aux: int, flags: int, d_verdicts: int, stream: int = 0):
        self._ck(self.L.clk_classify_hdr16_dev(self.h, d_recs, n, d_aux or None, flags, d_verdicts, stream or None))

    def classify_events(self, events: np.ndarray, flags: int = 0) -> np.ndarray:
        events = np.ascontiguousarray(events, dtype=EVENT_DTYPE)
        verd = np.empty(events.size, dtype=np.uint32)
        self._ck(self.L.clk_classify_events(self.h, events.ctypes.data, events.size, flags, verd.ctypes.data))
        return verd

    def dns_answers(self, blob: np.ndarray, offs16: np.ndarray, verdicts: np.ndarray, apply: bool = False, now: int = 0) -> np.ndarray:
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        offs16 = np.ascontiguousarray(offs16, dtype=np.uint32)
        verdicts = np.ascontiguousarray(verdicts, dtype=np.uint32)
        n = offs16.size - 1
        cap = int(((verdicts >> 27) & 31)[((verdicts >> 24) & 7) == 0].sum())
        out = np.zeros(cap, dtype=DNSUPD_DTYPE)
        k = C.c_uint64()
        self._ck(self.L.clk_dns_answers(self.h, blob.ctypes.data, offs16.ctypes.data, n, verdicts.ctypes.data,
                                        out.ctypes.data, cap, C.byref(k), 1 if apply else 0, now))
        return out[:k.value]

    # ---- counters -----------------------------------------------------------------------------------
    def counters(self, merged: bool = False) -> np.ndarray:
        """metrics_map rows: this GPU's table, or (merged=True) the all-rank view of the last counters_import."""
        fn = self.L.clk_counters_read_global if merged else self.L.clk_counters_read
        n = C.c_uint32()
        self._ck(fn(self.h, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=METRIC_DTYPE)
        if n.value:
            self._ck(fn(self.h, out.ctypes.data, n.value, C.byref(n)))
        return out

    def counters_reset(self):
        self._ck(self.L.clk_counters_reset(self.h))

    def counters_block_rows(self) -> int:
        n = C.c_uint32()
        self._ck(self.L.clk_counters_block_rows(self.h, C.byref(n)))
        return n.value

    def counters_export(self, d_block: int, stream: int = 0):
        self._ck(self.L.clk_counters_export(self.h, d_block, stream or None))

    def counters_import(self, d_blocks: int, n_blocks: int, stream: int = 0):
        self._ck(self.L.clk_counters_import(self.h, d_blocks, n_blocks, stream or None))

    # per-batch form of the collective: dense per-container totals (rows x 32 B)
    def totals_rows(self) -> int:
        n = C.c_uint32()
        self._ck(self.L.clk_totals_rows(self.h, C.byref(n)))
        return n.value

    def totals_export(self, d_block: int, stream: int = 0):
        self._ck(self.L.clk_totals_export(self.h, d_block, stream or None))

    def totals_merge(self, d_blocks: int, n_blocks: int) -> np.ndarray:
        """Sum gathered clk_total blocks by cgroup_id (synchronises the device); rows sorted by cgroup_id."""
        out = np.zeros(n_blocks * self.totals_rows(), dtype=abi.TOTAL_DTYPE)
        n = C.c_uint32()
        self._ck(self.L.clk_totals_merge(self.h, d_blocks, n_blocks, out.ctypes.data, out.shape[0], C.byref(n)))
        return out[:n.value]

    # multi-GPU: the collective behind the C ABI (NCCL is dlopen()ed by the library)
    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = C.create_string_buffer(unique_id, len(unique_id))
        self._ck(self.L.clk_comm_init(self.h, buf, len(unique_id), rank, world))

    def comm_destroy(self):
        self._ck(self.L.clk_comm_destroy(self.h))

    def allgather_counters(self, stream: int = 0):
        """Per batch, enqueue only: per-container totals of this GPU -> all-gather of every GPU's block."""
        self._ck(self.L.clk_allgather_counters(self.h, stream or None))

    def totals_global(self) -> np.ndarray:
        """The gathered totals summed by cgroup_id (synchronises the device)."""
        info_w = C.c_int()
        self._ck(self.L.clk_comm_info(self.h, None, C.byref(info_w)))
        out = np.zeros(max(1, info_w.value) * self.totals_rows(), dtype=abi.TOTAL_DTYPE)
        n = C.c_uint32()
        self._ck(self.L.clk_totals_read_global(self.h, out.ctypes.data, out.shape[0], C.byref(n)))
        return out[:n.value]

    def allgather_metrics(self, stream: int = 0):
        """On read, enqueue only: per-key export -> all-gather -> import into the merged view (counters(merged=True))."""
        self._ck(self.L.clk_allgather_metrics(self.h, stream or None))

    def stats(self) -> Stats:
        s = Stats()
        self._ck(self.L.clk_get_stats(self.h, C.byref(s)))
        return s

    def sync(self):
        self._ck(self.L.clk_sync(self.h))


def pack_hdr16(pkts: np.ndarray, stride: int = 64, pinned: bool = False):
    """clk_pack_hdr16: slot-layout frames -> (records[n] of abi.HDR16_DTYPE, aux bytes [n_aux * 64])."""
    pkts = np.ascontiguousarray(pkts, dtype=np.uint8)
    n = pkts.size // stride
    alloc = (lambda nb: pinned_empty(nb)) if pinned else (lambda nb: np.empty(nb, dtype=np.uint8))
    recs = alloc(max(1, n) * 16)[: n * 16].view(abi.HDR16_DTYPE)
    cap = max(1, n // 8 + 1024)
    while True:
        aux, n_aux = alloc(cap * 64), C.c_uint64()
        rc = lib().clk_pack_hdr16(pkts.ctypes.data, stride, n, recs.ctypes.data, aux.ctypes.data, cap, C.byref(n_aux))
        if rc == 0:
            return recs, aux[: n_aux.value * 64]
        if rc != abi.E2BIG or cap >= n:
            raise ClkError(rc, lib().clk_last_error(None).decode())
        cap = n


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0 calls this and hands the bytes to the other ranks)."""
    buf, n = C.create_string_buffer(256), C.c_size_t()
    rc = lib().clk_comm_unique_id(buf, 256, C.byref(n))
    if rc:
        raise ClkError(rc, lib().clk_last_error(None).decode())
    return buf.raw[:n.value]


def domain_hash(s: str) -> int:
    b = s.encode()
    return lib().clk_domain_hash(b, len(b))


def validate_dst(s: str) -> bool:
    return lib().clk_validate_dst(s.encode()) == 0


def ip_to_u32(s: str) -> int:
    v = C.c_uint32()
    lib().clk_ip_to_u32(s.encode(), C.byref(v))
    return v.value


def cidr_to_addr_mask(s: str):
    a, m = C.c_uint32(), C.c_uint32()
    if lib().clk_cidr_to_addr_mask(s.encode(), C.byref(a), C.byref(m)):
        raise ValueError(s)
    return a.value, m.value
