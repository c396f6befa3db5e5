"""The C-ABI library: loads, exports every symbol include/clawker_gpu.h declares, and its BPF-map-shaped
ops behave like the cilium/ebpf calls they replace (no compute calls: this file runs without a GPU)."""
import ctypes as C
import re
from pathlib import Path

import pytest

from clawker_b200 import abi, ebpf, engine

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    hdr = (ROOT / "include" / "clawker_gpu.h").read_text()
    declared = set(re.findall(r"^(?:int|void|uint32_t|const char \*)\s*(clk_[a-z0-9_]+)\(", hdr, flags=re.M))
    L = engine.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert declared == set(engine.EXPORTS), sorted(declared ^ set(engine.EXPORTS))
    assert L.clk_abi_version() == 3


def test_struct_sizes_match_reference_layouts():
    # internal/ebpf/bpf/common.h:80-113,163-169
    assert C.sizeof(abi.ContainerConfig) == 28 and C.sizeof(abi.DNSEntry) == 8
    assert C.sizeof(abi.RouteKey) == 8 and C.sizeof(abi.RouteVal) == 4 and C.sizeof(abi.MetricKey) == 16


@pytest.fixture()
def eng():
    e = engine.Engine(host_only=True, max_dns_cache=4, max_routes=2, max_containers=2)
    yield e
    e.close()


def test_map_semantics(eng):
    k = C.c_uint32(ebpf.IPToUint32("140.82.121.4"))
    assert eng.map_lookup(abi.MAP_DNS_CACHE, k, abi.DNSEntry()) is None
    eng.map_update(abi.MAP_DNS_CACHE, k, abi.DNSEntry(7, 100))
    eng.map_update(abi.MAP_DNS_CACHE, k, abi.DNSEntry(9, 200))                       # UpdateAny overwrites
    v = eng.map_lookup(abi.MAP_DNS_CACHE, k, abi.DNSEntry())
    assert (v.domain_hash, v.expire_ts) == (9, 200) and eng.map_count(abi.MAP_DNS_CACHE) == 1
    assert eng.map_delete(abi.MAP_DNS_CACHE, k) == 0
    assert eng.map_delete(abi.MAP_DNS_CACHE, k) == abi.ENOENT                        # ebpf.ErrKeyNotExist
    for i in range(4):
        eng.map_update(abi.MAP_DNS_CACHE, C.c_uint32(i + 1), abi.DNSEntry(i, 0))
    with pytest.raises(engine.ClkError) as ei:                                       # max_entries reached: E2BIG
        eng.map_update(abi.MAP_DNS_CACHE, C.c_uint32(99), abi.DNSEntry(1, 0))
    assert ei.value.code == abi.E2BIG
    eng.map_update(abi.MAP_DNS_CACHE, C.c_uint32(1), abi.DNSEntry(42, 0))            # existing key still updatable
    with pytest.raises(engine.ClkError):                                             # wrong value size
        eng._ck(eng.L.clk_map_update(eng.h, abi.MAP_DNS_CACHE, C.byref(k), 4, C.byref(C.c_uint8(1)), 1))
    with pytest.raises(engine.ClkError):                                             # not a CIDR mask
        eng.map_update(abi.MAP_CIDR, abi.Cidr(0, 0x00FF00FF), None)


def test_dns_update_and_gc(eng):
    m = ebpf.Manager(engine=eng)
    m.UpdateDNSCache(ebpf.IPToUint32("1.1.1.1"), 5, 300, now=1000)                   # expire 1300
    m.UpdateDNSCache(ebpf.IPToUint32("2.2.2.2"), 6, 60, now=1000)                    # expire 1060
    assert m.GarbageCollectDNS(now=1060) == 0                                        # expire_ts < now is strict (manager.go:450)
    assert m.GarbageCollectDNS(now=1061) == 1
    assert m.GarbageCollectDNS(now=1061) == 0
    assert eng.map_count(abi.MAP_DNS_CACHE) == 1


def test_manager_enable_disable_bypass(eng):
    m = ebpf.Manager(engine=eng)
    cfg = ebpf.NewContainerConfig("172.20.0.200", "172.20.0.201", "172.20.0.1", "172.20.0.0/16", "", 0, 10000)
    assert cfg.host_proxy_ip == 0 and cfg.net_mask == ebpf.IPToUint32("255.255.0.0")
    with pytest.raises(ValueError):
        ebpf.NewContainerConfig("a", "b", "c", "not-a-cidr", "", 0, 0)
    m.Bypass(7)
    m.Enable(7, "/sys/fs/cgroup/x", cfg)                                             # Enable clears a stale bypass flag
    assert eng.map_count(abi.MAP_BYPASS) == 0 and m.LookupContainer(7).egress_port == 10000
    m.Unbypass(7)                                                                    # missing key is success
    m.Disable(7)
    m.Disable(7)                                                                     # idempotent
    with pytest.raises(ebpf.KeyNotExist):
        m.LookupContainer(7)
    m.SyncRoutes([ebpf.Route(1, 22, 10001), ebpf.Route(2, 22, 10002)])
    m.SyncRoutes([ebpf.Route(3, 443, 10001)])                                        # replace, not merge
    assert eng.map_count(abi.MAP_ROUTE) == 1
    with pytest.raises(engine.ClkError):
        m.SyncRoutes([ebpf.Route(i, 1, 2) for i in range(3)])                        # > max_entries


def test_delete_errors_other_than_missing_key_surface(eng):
    """TestClearBypass_IgnoresMissingEntry / _WrapsOtherErrors and TestDeleteExpiredDNSEntries_EmptyReturnsZero
    (internal/ebpf/manager_test.go:216-270, 344-354): a missing key is success, anything else is an error that is NOT the
    key-not-exist code — so a caller cannot mistake a failed clear for a cleared flag."""
    m = ebpf.Manager(engine=eng)
    assert eng.map_delete(abi.MAP_BYPASS, C.c_uint64(456)) == abi.ENOENT             # never set: the common case, silent
    m.Unbypass(456)
    with pytest.raises(engine.ClkError) as ei:                                       # a key of the wrong size (EINVAL from the map)
        eng._ck(eng.L.clk_map_delete(eng.h, abi.MAP_BYPASS, C.byref(C.c_uint32(1)), 4), ok=(0, abi.ENOENT))
    assert ei.value.code == abi.EINVAL and ei.value.code != abi.ENOENT
    with pytest.raises(engine.ClkError) as ei:                                       # no such map
        eng._ck(eng.L.clk_map_delete(eng.h, 99, C.byref(C.c_uint64(1)), 8), ok=(0, abi.ENOENT))
    assert ei.value.code != abi.ENOENT
    assert m.GarbageCollectDNS(now=10**9) == 0                                       # nothing expired, nothing touched


def test_no_cpu_fallback(eng):
    import numpy as np
    for call in (lambda: eng.publish(), lambda: eng.classify_slots(np.zeros((4, 64), np.uint8)),
                 lambda: eng.classify_events(np.zeros(1, abi.EVENT_DTYPE)), lambda: eng.counters()):
        with pytest.raises(engine.ClkError) as ei:
            call()
        assert ei.value.code == abi.ENOSYS


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under clawker_b200/ (or the C++ sources) may reference it."""
    for p in (ROOT / "clawker_b200").rglob("*"):
        if p.suffix in (".py", ".cpp", ".cu", ".cuh", ".h") and p.is_file():
            txt = p.read_text()
            assert "oracle" not in txt.replace("test oracle", "").lower() or p.name == "capture.py", p


def test_graft_entry_build_runs():
    """The driver's "does it build" check: every extension compiles for sm_100a and every declared symbol resolves."""
    import __graft_entry__ as g
    g.build()
    assert engine.LIB_PATH.exists() and engine.CLI_PATH.exists()
