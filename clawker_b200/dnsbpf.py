"""Python mirror of the reference's CoreDNS plugin internal/dnsbpf (dnsbpf.go:32-105, bpfmap.go:39-48).

`Handler.write_a_records` is the per-response path the plugin runs inside CoreDNS; the batch form of the
same function over captured DNS-answer packets is the GPU kernel behind Engine.dns_answers (clk_dns_answers).
"""
from __future__ import annotations

from dataclasses import dataclass, field

from .engine import Engine, domain_hash, ip_to_u32

PLUGIN_NAME = "dnsbpf"                       # setup.go:12
DEFAULT_PIN_PATH = "/sys/fs/bpf/clawker/dns_cache"   # bpfmap.go:13 (interface parity only)
TYPE_A, TYPE_AAAA, TYPE_CNAME = 1, 28, 5


@dataclass
class RR:                                     # the part of dns.RR the plugin reads
    name: str
    rrtype: int
    ttl: int
    a: str = ""                               # dotted IPv4 for A records


@dataclass
class Msg:                                    # the part of dns.Msg the plugin reads
    answer: list = field(default_factory=list)


class MapWriter:                              # dnsbpf.go:32-34
    def Update(self, ip: int, domain_hash: int, ttl_seconds: int):
        raise NotImplementedError


class EngineMap(MapWriter):
    """BPFMap (bpfmap.go:24-48) backed by the GPU engine's dns_cache."""

    def __init__(self, engine: Engine, now: int = 0):
        self.e, self.now = engine, now

    def Update(self, ip: int, domain_hash: int, ttl_seconds: int):
        self.e.dns_update(ip, domain_hash, ttl_seconds, self.now)


def zone_to_domain(zone: str) -> str:
    """dnsbpf.go:103-105: strings.TrimSuffix(zone, ".")."""
    return zone[:-1] if zone.endswith(".") else zone


class Handler:
    """dnsbpf.Handler (dnsbpf.go:38-70)."""

    def __init__(self, zone: str, map_writer: MapWriter | None, next_handler=None):
        self.Zone, self.Map, self.Next = zone, map_writer, next_handler

    def Name(self) -> str:
        return PLUGIN_NAME

    def ServeDNS(self, request):
        """Capture the next plugin's response, write its A records, pass the response on unchanged."""
        msg = self.Next(request) if self.Next else None
        if msg is None:
            return None
        if self.Map is not None:
            self.write_a_records(msg)
        return msg

    def write_a_records(self, msg: Msg):
        """dnsbpf.go:75-97."""
        h = domain_hash(zone_to_domain(self.Zone))
        for rr in msg.answer:
            if rr.rrtype != TYPE_A:
                continue
            ip = ip_to_u32(rr.a)
            if ip == 0:
                continue
            self.Map.Update(ip, h, rr.ttl if rr.ttl != 0 else 60)
