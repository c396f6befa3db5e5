"""Python mirror of the rule-facing part of the reference's internal/firewall package.

Only what defines the allowlist tables: config.EgressRule, ValidateDst, and the three
generators' table outputs.  All of it is computed by the C++ rule compiler inside libclawker_gpu.so
(clk_rules_compile / clk_rules_dump); nothing here re-implements it.
"""
from __future__ import annotations

from dataclasses import dataclass

from .engine import Engine, validate_dst as _validate_dst

ENVOY_EGRESS_PORT = 10000        # internal/config/consts.go:110
ENVOY_TCP_PORT_BASE = 10001      # consts.go:112


@dataclass
class EgressRule:                # internal/config/schema.go:131-138
    dst: str
    proto: str = ""
    port: int = 0
    action: str = ""


def ValidateDst(dst: str) -> None:
    """firewall.ValidateDst (rules.go:33-107): raises ValueError when the reference returns an error."""
    if not _validate_dst(dst):
        raise ValueError(f"invalid destination {dst!r}")


class Compiled:
    def __init__(self, zones, server_names, tcp_mappings, warnings):
        self.zones, self.server_names, self.tcp_mappings, self.warnings = zones, server_names, tcp_mappings, warnings


def compile_rules(engine: Engine, rules, egress_port: int = ENVOY_EGRESS_PORT, tcp_port_base: int = ENVOY_TCP_PORT_BASE) -> Compiled:
    """normalizeAndDedup + GenerateCorefile zones + serverNames + TCPMappings + syncRoutes
    (rules.go:138-161, coredns.go:26-96, envoy.go:54-83,905-915, manager.go:1356-1377) into the engine's tables."""
    w = engine.compile_rules(rules, egress_port, tcp_port_base)
    maps = []
    for line in engine.rules_dump(2):
        dst, port, eport = line.rsplit(":", 2)
        maps.append((dst, int(port), int(eport)))
    return Compiled(engine.rules_dump(0), engine.rules_dump(1), maps, w)
