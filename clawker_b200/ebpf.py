"""Python mirror of the reference's Go package internal/ebpf (the map manager), backed by the GPU engine.

Same names, argument meaning and error behaviour as internal/ebpf/manager.go:31-470 and types.go:28-125,
so tests read like internal/ebpf/manager_test.go.  What attaches BPF programs to a cgroup in the reference
(link.AttachCgroup, manager.go:252-292) has no counterpart: enforcement here is "the cgroup has a
container_map row", and packets reach the classifier through clk_classify_*.
"""
from __future__ import annotations

import ctypes as C
import ipaddress

from . import abi
from .abi import ContainerConfig, DNSEntry, Endpoint
from .engine import ClkError, Engine, cidr_to_addr_mask, domain_hash as _domain_hash, ip_to_u32

PIN_PATH = "/sys/fs/bpf/clawker"            # types.go:28 (kept for interface parity; nothing is pinned)
ACTION_ALLOW, ACTION_DENY, ACTION_BYPASS = 0, 1, 2      # types.go:76-80


class KeyNotExist(KeyError):
    """ebpf.ErrKeyNotExist"""


def IPToUint32(ip: str) -> int:               # types.go:86-92
    return ip_to_u32(ip)


def Uint32ToIP(n: int) -> str:                # types.go:95-99
    return str(ipaddress.IPv4Address(int(n).to_bytes(4, "little")))


def CIDRToAddrMask(cidr: str):                # types.go:101-111 (raises ValueError on a bad CIDR)
    return cidr_to_addr_mask(cidr)


def DomainHash(domain: str) -> int:           # types.go:121-125
    return _domain_hash(domain)


def NewContainerConfig(envoy_ip: str, coredns_ip: str, gateway_ip: str, cidr: str, host_proxy_ip: str,
                       host_proxy_port: int, egress_port: int) -> ContainerConfig:
    """manager.go:473-494."""
    try:
        net_addr, net_mask = CIDRToAddrMask(cidr)
    except ValueError as ex:
        raise ValueError(f"parsing CIDR {cidr}: {ex}") from None
    cfg = ContainerConfig(IPToUint32(envoy_ip), IPToUint32(coredns_ip), IPToUint32(gateway_ip), net_addr, net_mask,
                          0, host_proxy_port, egress_port)
    if host_proxy_ip != "":
        cfg.host_proxy_ip = IPToUint32(host_proxy_ip)
    return cfg


class Route:                                  # manager.go:466-470
    def __init__(self, domain_hash: int, dst_port: int, envoy_port: int):
        self.DomainHash, self.DstPort, self.EnvoyPort = domain_hash, dst_port, envoy_port


class Manager:
    """ebpf.Manager (manager.go:31-463)."""

    def __init__(self, engine: Engine | None = None, **engine_kw):
        self.e = engine if engine is not None else Engine(**engine_kw)

    # Load / OpenPinned (manager.go:41-133): the engine already holds the (empty) maps; with a pin path they are
    # shared with the `ebpf-manager` verbs through the state file (the bpffs analogue, clk_state_save / _load).
    def Load(self, pin_path: str | None = None):
        if pin_path:
            self.e.state_save(pin_path)

    def OpenPinned(self, pin_path: str | None = None):
        if pin_path:
            self.e.state_load(pin_path)

    def Close(self):
        self.e.close()

    def Enable(self, cgroup_id: int, cgroup_path: str, cfg: ContainerConfig):
        """manager.go:236-298: clear a stale bypass flag, then container_map[cgroup] = cfg."""
        self.e.map_delete(abi.MAP_BYPASS, C.c_uint64(cgroup_id))                 # ErrKeyNotExist is success (clearBypass :222-231)
        try:
            self.e.map_update(abi.MAP_CONTAINER, C.c_uint64(cgroup_id), cfg)
        except ClkError as ex:
            raise RuntimeError(f"ebpf enable: updating container_map: {ex}") from None

    def Disable(self, cgroup_id: int):
        """manager.go:301-337: delete the container row and any bypass flag; missing keys are fine."""
        self.e.map_delete(abi.MAP_CONTAINER, C.c_uint64(cgroup_id))
        self.e.map_delete(abi.MAP_BYPASS, C.c_uint64(cgroup_id))

    def SyncRoutes(self, routes):
        """manager.go:347-386, as one atomic replace."""
        self.e.routes_replace([(r.DomainHash, r.DstPort, r.EnvoyPort) for r in routes])

    def Bypass(self, cgroup_id: int):                                            # manager.go:389-396
        self.e.map_update(abi.MAP_BYPASS, C.c_uint64(cgroup_id), C.c_uint8(1))

    def Unbypass(self, cgroup_id: int):                                          # manager.go:399-405
        self.e.map_delete(abi.MAP_BYPASS, C.c_uint64(cgroup_id))

    def UpdateDNSCache(self, ip: int, domain_hash: int, ttl_seconds: int, now: int = 0):     # manager.go:408-414
        self.e.dns_update(ip, domain_hash, ttl_seconds, now)

    def GarbageCollectDNS(self, now: int = 0) -> int:                            # manager.go:443-456
        return self.e.dns_gc(now)

    def LookupContainer(self, cgroup_id: int) -> ContainerConfig:               # manager.go:459-463
        v = self.e.map_lookup(abi.MAP_CONTAINER, C.c_uint64(cgroup_id), ContainerConfig())
        if v is None:
            raise KeyNotExist(cgroup_id)
        return v

    def ScrapeMetrics(self, merged: bool = False):
        """The reader the reference planned (outstanding-features.md:262-300): the metrics_map rows (see exporter.py)."""
        return self.e.counters(merged=merged)

    # ---- extension: the packet -> socket identity binding -----------------------------
    def BindEndpoint(self, ip: str, cgroup_id: int, uid: int = 1000, so_mark: int = 0):
        self.e.map_update(abi.MAP_ENDPOINT, C.c_uint32(IPToUint32(ip)), Endpoint(cgroup_id, uid, so_mark))

    def Publish(self):
        self.e.publish()
