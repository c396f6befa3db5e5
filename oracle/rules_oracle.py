"""Python restatement of clawker's Go rule compiler (TEST INFRASTRUCTURE — see oracle/README.md).

Each function cites the reference code it follows; it is pinned by the reference's own unit-test
vectors transcribed as data under tests/golden/ (tests/test_rules_oracle.py).  The product's rule
compiler is C++ inside libclawker_gpu.so (clk_rules_compile); this module is only its checker.
"""
from __future__ import annotations

import ipaddress
import re
from dataclasses import dataclass, field

# internal/firewall/coredns.go:35-42
INTERNAL_HOSTS = ["docker.internal", "otel-collector", "jaeger", "prometheus", "loki", "grafana"]


@dataclass
class EgressRule:          # internal/config/schema.go:124-138
    dst: str
    proto: str = ""
    port: int = 0
    action: str = ""
    path_rules: list = field(default_factory=list)     # [(path, action), ...]   PathRule schema.go:124-127
    path_default: str = ""


def domain_hash(domain: str) -> int:
    """ebpf.DomainHash, internal/ebpf/types.go:121-125 (FNV-1a-32 of strings.ToLower)."""
    h = 0x811C9DC5
    for c in domain.lower().encode():
        h = ((h ^ c) * 0x01000193) & 0xFFFFFFFF
    return h


def ip_to_u32(s: str) -> int:
    """ebpf.IPToUint32, types.go:86-92: 4 network-order bytes read as a native (little-endian) u32."""
    try:
        b = ipaddress.IPv4Address(s).packed
    except ValueError:
        return 0
    return int.from_bytes(b, "little")


def _parse_ip(s: str):
    """net.ParseIP: dotted IPv4 (no leading zeros, Go >= 1.17) or IPv6 text."""
    try:
        return ipaddress.ip_address(s)
    except ValueError:
        return None


def _parse_cidr(s: str):
    """net.ParseCIDR: addr '/' decimal prefix; returns the masked network."""
    if "/" not in s:
        return None
    a, _, p = s.partition("/")
    ip = _parse_ip(a)
    if ip is None or not re.fullmatch(r"[0-9]{1,3}", p):
        return None
    if len(p) > 1 and p[0] == "0":      # Go's dtoi accepts "08"? ParseCIDR uses dtoi: leading zeros ok
        pass
    n = int(p)
    if n > (32 if ip.version == 4 else 128):
        return None
    return ipaddress.ip_network((ip, n), strict=False)


def cidr_to_addr_mask(cidr: str):
    """ebpf.CIDRToAddrMask, types.go:101-111 (IPv4)."""
    net = _parse_cidr(cidr)
    if net is None or net.version != 4:
        raise ValueError(cidr)
    return int.from_bytes(net.network_address.packed, "little"), int.from_bytes(net.netmask.packed, "little")


def is_ip_or_cidr(s: str) -> bool:
    """internal/firewall/coredns.go:109-115."""
    return _parse_ip(s) is not None or _parse_cidr(s) is not None


def normalize_domain(d: str) -> str:
    """coredns.go:119-122: strip ONE leading dot and ONE trailing dot."""
    if d.startswith("."):
        d = d[1:]
    if d.endswith("."):
        d = d[:-1]
    return d


def is_wildcard_domain(d: str) -> bool:
    """coredns.go:126-128."""
    return d.startswith(".")


def is_allow_domain(r: EgressRule) -> bool:
    """coredns.go:98-106."""
    a = r.action.lower()
    if a not in ("allow", ""):
        return False
    return not is_ip_or_cidr(r.dst)


def validate_dst(dst: str) -> bool:
    """firewall.ValidateDst, internal/firewall/rules.go:33-107.  True = accepted."""
    if dst == "":
        return False
    n = normalize_domain(dst)
    if n == "":
        return False
    if is_ip_or_cidr(n):
        return not dst.startswith(".")
    if len(n) > 253:
        return False
    last = "."
    non_numeric = False
    partlen = 0
    for c in n:
        if "a" <= c <= "z" or c == "_":
            non_numeric = True
            partlen += 1
        elif "0" <= c <= "9":
            partlen += 1
        elif c == "-":
            if last == ".":
                return False
            non_numeric = True
            partlen += 1
        elif c == ".":
            if last in ".-":
                return False
            if partlen > 63:
                return False
            partlen = 0
        else:
            return False
        last = c
    if last in ".-" or partlen > 63:
        return False
    return non_numeric


def normalize_rule(r: EgressRule) -> EgressRule:
    """rules.go:113-125."""
    proto = r.proto or "tls"
    action = r.action or "allow"
    port = r.port
    if port == 0 and proto.lower() == "tls":
        port = 443
    return EgressRule(r.dst, proto, port, action, list(r.path_rules), r.path_default)


def normalize_and_dedup(rules):
    """rules.go:138-161.  Returns (rules, n_warnings)."""
    seen = set()
    out = []
    warnings = 0
    for r in rules:
        r = normalize_rule(r)
        if normalize_domain(r.dst) == "":
            warnings += 1
            continue
        key = f"{r.dst}:{r.proto}:{r.port}"
        if key in seen:
            continue
        seen.add(key)
        out.append(r)
    return out, warnings


def corefile_zones(rules):
    """Zone set of GenerateCorefile, coredns.go:26-96: allow-domain rules in order, normalised,
    deduped, reserved internal hosts skipped; then the six internal hosts."""
    reserved = set(INTERNAL_HOSTS)
    emitted = set()
    domains = []
    for r in rules:
        if not is_allow_domain(r):
            continue
        d = normalize_domain(r.dst)
        if d in reserved or d in emitted:
            continue
        emitted.add(d)
        domains.append(d)
    return domains + list(INTERNAL_HOSTS)


def tcp_default_port(r: EgressRule) -> int:
    """envoy.go:671-679."""
    if r.port != 0:
        return r.port
    return 22 if r.proto.lower() == "ssh" else 443


def tcp_mappings(rules, tcp_port_base: int = 10001):
    """firewall.TCPMappings, envoy.go:54-83 -> [(dst, dst_port, envoy_port)]."""
    out = []
    idx = 0
    for r in rules:
        if r.action.lower() not in ("allow", ""):
            continue
        if is_ip_or_cidr(r.dst):
            continue
        if r.proto.lower() not in ("ssh", "tcp"):
            continue
        out.append((normalize_domain(r.dst), tcp_default_port(r), tcp_port_base + idx))
        idx += 1
    return out


def server_names(dst: str, exact_domains) -> list:
    """envoy.go:905-915."""
    d = normalize_domain(dst)
    if is_wildcard_domain(dst):
        if d in exact_domains:
            return ["." + d]
        return ["." + d, d]
    return [d]


def http_domains(dst: str, exact_domains) -> list:
    """envoy.go:932-944."""
    d = normalize_domain(dst)
    if is_wildcard_domain(dst):
        if d in exact_domains:
            return ["*." + d, "*." + d + ":*"]
        return ["*." + d, "*." + d + ":*", d, d + ":*"]
    return [d, d + ":*"]


def http_rules(rules):
    """Rule classification of GenerateEnvoyConfig, envoy.go:176-209: what lands in httpRules (port 0 -> 80)."""
    out = []
    for r in rules:
        if r.action.lower() not in ("allow", "") or is_ip_or_cidr(r.dst) or r.proto.lower() != "http":
            continue
        out.append(EgressRule(r.dst, r.proto, r.port or 80, r.action, list(r.path_rules), r.path_default))
    return out


def http_routes(r: EgressRule) -> list:
    """buildHTTPFilterChain / buildHTTPRoutes, envoy.go:403-425,566-607: [(prefix, allow)] in match order."""
    if not r.path_rules:
        return [("/", True)]
    routes = [(p, a.lower() == "allow") for p, a in r.path_rules]          # strings.EqualFold(pr.Action, "allow")
    d = r.path_default.lower()
    routes.append(("/", not (d == "" or d == "deny")))
    return routes


def http_vhosts(rules) -> list:
    """http_egress_routes virtual hosts in config order (the trailing deny_all host "*" is implicit):
    [(domains, routes)], envoy.go:403-470 with httpExactDomains of :222-227."""
    hs = http_rules(rules)
    exact = {normalize_domain(r.dst) for r in hs if not is_wildcard_domain(r.dst)}
    return [(http_domains(r.dst, exact), http_routes(r)) for r in hs]


def tls_rules(rules):
    """Rule classification of GenerateEnvoyConfig, envoy.go:176-209: what lands in tlsRules."""
    out = []
    for r in rules:
        if r.action.lower() not in ("allow", ""):
            continue
        if is_ip_or_cidr(r.dst):
            continue
        p = r.proto.lower()
        if p in ("ssh", "tcp", "http"):
            continue
        out.append(r)
    return out


def sni_table(rules) -> list:
    """All filter_chain_match.server_names of the egress listener in chain order
    (buildEgressListener envoy.go:358-393 + serverNames), duplicates removed keeping first."""
    tls = tls_rules(rules)
    exact = {normalize_domain(r.dst) for r in tls if not is_wildcard_domain(r.dst)}   # envoy.go:215-220
    out = []
    for r in tls:
        for s in server_names(r.dst, exact):
            if s not in out:
                out.append(s)
    return out


def cidr_rules(rules) -> list:
    """A12 (extension): allow rules whose dst is an IPv4 address or CIDR -> (addr, mask)."""
    out = []
    for r in rules:
        if r.action.lower() not in ("allow", ""):
            continue
        ip = _parse_ip(r.dst)
        if ip is not None:
            if ip.version == 4:
                out.append((int.from_bytes(ip.packed, "little"), 0xFFFFFFFF))
            continue
        net = _parse_cidr(r.dst)
        if net is not None and net.version == 4:
            out.append((int.from_bytes(net.network_address.packed, "little"),
                        int.from_bytes(net.netmask.packed, "little")))
    return out


@dataclass
class Compiled:
    zones: list = field(default_factory=list)
    snis: list = field(default_factory=list)
    routes: dict = field(default_factory=dict)     # (domain_hash, dst_port) -> envoy_port
    mappings: list = field(default_factory=list)
    cidrs: list = field(default_factory=list)
    http: list = field(default_factory=list)       # [(domains, [(prefix, allow)])]
    warnings: int = 0


def compile_rules(rules, tcp_port_base: int = 10001) -> Compiled:
    """ensureConfigs + syncRoutes, internal/firewall/manager.go:753-836,1356-1377:
    normalizeAndDedup -> Corefile zones, Envoy server_names, route_map rows."""
    rules = [r if isinstance(r, EgressRule) else EgressRule(**r) for r in rules]
    for r in rules:                                # path rules as (path, action) pairs or {"path":, "action":} dicts
        r.path_rules = [(x["path"], x.get("action", "")) if isinstance(x, dict) else (x[0], x[1]) for x in r.path_rules]
    rs, w = normalize_and_dedup(rules)
    c = Compiled(warnings=w)
    c.zones = corefile_zones(rs)
    c.snis = sni_table(rs)
    c.mappings = tcp_mappings(rs, tcp_port_base)
    for dst, port, eport in c.mappings:            # later duplicates overwrite (map Update, manager.go:1367-1374)
        c.routes[(domain_hash(dst), port)] = eport
    c.cidrs = cidr_rules(rs)
    c.http = http_vhosts(rs)
    return c
