"""The rule compiler against the reference's own unit-test vectors (tests/golden/reference_vectors.py):
once for the oracle's Python restatement, once for the product's C++ compiler behind the C-ABI."""
import pytest

from clawker_b200 import dnsbpf, ebpf, engine, firewall
from oracle import rules_oracle as ro

from golden import reference_vectors as V


def R(t):
    return ro.EgressRule(*t)


# ---- oracle restatement ------------------------------------------------------------------------------
@pytest.mark.parametrize("dst,want_err", V.VALIDATE_DST)
def test_oracle_validate_dst(dst, want_err):
    assert ro.validate_dst(dst) == (not want_err)


@pytest.mark.parametrize("name,rules,want", V.TCP_MAPPINGS)
def test_oracle_tcp_mappings(name, rules, want):
    assert ro.tcp_mappings([R(r) for r in rules], 10001) == want


def test_oracle_corefile_zones():
    assert ro.corefile_zones([R(r) for r in V.COREFILE_RULES]) == V.COREFILE_ZONES
    z = ro.corefile_zones([R(r) for r in V.COREFILE_WILDCARD_RULES])
    assert z[:2] == V.COREFILE_WILDCARD_ZONES_PREFIX and ".datadoghq.com" not in z


def test_oracle_server_names_http_domains():
    for dst, exact, want in V.SERVER_NAMES:
        assert ro.server_names(dst, set(exact)) == want
    for dst, exact, want in V.HTTP_DOMAINS:
        assert ro.http_domains(dst, set(exact)) == want


def test_oracle_normalize_and_dedup():
    for rules, want, warns in V.NORMALIZE_DEDUP:
        out, w = ro.normalize_and_dedup([R(r) for r in rules])
        assert [(r.dst, r.port) for r in out] == want and w == warns
    for d, want in V.IS_WILDCARD:
        assert ro.is_wildcard_domain(d) == want
    for d, want in V.NORMALIZE_DOMAIN:
        assert ro.normalize_domain(d) == want


def test_oracle_domain_hash():
    for variants, ref in V.DOMAIN_HASH_EQUAL:
        assert {ro.domain_hash(v) for v in variants} == {ro.domain_hash(ref)}
    for s, h in V.DOMAIN_HASH_KAT:
        assert ro.domain_hash(s) == h
    m = ro.tcp_mappings([R(V.WILDCARD_ROUTE_RULE)])
    assert ro.domain_hash(m[0][0]) == ro.domain_hash(ro.normalize_domain(".example.com"))
    for ip, want in V.IP_TO_U32:
        assert ro.ip_to_u32(ip) == want


# ---- product (C++ behind the C-ABI; host-only engine: no GPU needed for the control plane) ------------
@pytest.fixture()
def eng():
    e = engine.Engine(host_only=True)
    yield e
    e.close()


@pytest.mark.parametrize("dst,want_err", V.VALIDATE_DST)
def test_product_validate_dst(dst, want_err):
    assert engine.validate_dst(dst) == (not want_err)
    if want_err:
        with pytest.raises(ValueError):
            firewall.ValidateDst(dst)
    else:
        firewall.ValidateDst(dst)


@pytest.mark.parametrize("name,rules,want", V.TCP_MAPPINGS)
def test_product_tcp_mappings(eng, name, rules, want):
    c = firewall.compile_rules(eng, [firewall.EgressRule(*r) for r in rules])
    assert c.tcp_mappings == want
    # syncRoutes: route_map rows == {DomainHash(dst), port} -> envoy_port  (manager.go:1356-1377)
    from clawker_b200 import abi
    assert eng.map_count(abi.MAP_ROUTE) == len({(d, p) for d, p, _ in want})
    for d, p, ep in want:
        v = eng.map_lookup(abi.MAP_ROUTE, abi.RouteKey(ebpf.DomainHash(d), p, 0), abi.RouteVal())
        assert v is not None and v.envoy_port == ep


def test_product_corefile_zones(eng):
    assert firewall.compile_rules(eng, [firewall.EgressRule(*r) for r in V.COREFILE_RULES]).zones == V.COREFILE_ZONES
    z = firewall.compile_rules(eng, [firewall.EgressRule(*r) for r in V.COREFILE_WILDCARD_RULES]).zones
    assert z[:2] == V.COREFILE_WILDCARD_ZONES_PREFIX and ".datadoghq.com" not in z


def test_product_server_names(eng):
    # serverNames() through the compiled SNI table: one tls rule per case (+ the exact sibling when given)
    for dst, exact, want in V.SERVER_NAMES:
        rules = [firewall.EgressRule(dst, "tls", 443, "allow")] + [firewall.EgressRule(x, "tls", 443, "allow") for x in exact]
        got = firewall.compile_rules(eng, rules).server_names
        assert got[:len(want)] == want and got[len(want):] == list(exact)


def test_product_normalize_and_dedup(eng):
    for rules, want, warns in V.NORMALIZE_DEDUP:
        c = firewall.compile_rules(eng, [firewall.EgressRule(*r) for r in rules])
        assert c.warnings == warns
        # every surviving tls rule shows up as a zone (normalised, deduped)
        zones = [z for z in c.zones if z not in ro.INTERNAL_HOSTS]
        seen = []
        for d, _ in want:
            n = ro.normalize_domain(d)
            if n not in seen:
                seen.append(n)
        assert zones == seen


def test_product_domain_hash_and_addresses():
    for variants, ref in V.DOMAIN_HASH_EQUAL:
        assert {ebpf.DomainHash(v) for v in variants} == {ebpf.DomainHash(ref)}
    for s, h in V.DOMAIN_HASH_KAT:
        assert ebpf.DomainHash(s) == h
    for ip, want in V.IP_TO_U32:
        assert ebpf.IPToUint32(ip) == want
    assert ebpf.Uint32ToIP(75059852) == "140.82.121.4"
    assert ebpf.CIDRToAddrMask("172.20.0.0/16") == (ebpf.IPToUint32("172.20.0.0"), ebpf.IPToUint32("255.255.0.0"))
    assert ebpf.CIDRToAddrMask("10.1.2.3/8")[0] == ebpf.IPToUint32("10.0.0.0")
    with pytest.raises(ValueError):
        ebpf.CIDRToAddrMask("10.0.0.0/33")


def test_product_equals_oracle_on_synthetic_rule_sets(eng):
    from clawker_b200 import capture
    for cfg in ("c1", "c2", "c3"):
        w = capture.make_workload(cfg, n_rules=600 if cfg == "c3" else None)
        c = firewall.compile_rules(eng, w.rules)
        o = ro.compile_rules(w.rules)
        assert c.zones == o.zones and c.server_names == o.snis and c.tcp_mappings == o.mappings and c.warnings == o.warnings


# ---- dnsbpf plugin surface ---------------------------------------------------------------------------
class RecordingMap(dnsbpf.MapWriter):        # dnsbpf_test.go:31-58 recordingMap
    def __init__(self):
        self.entries = []

    def Update(self, ip, domain_hash, ttl):
        self.entries.append((ip, domain_hash, ttl))


@pytest.mark.parametrize("zone,answers,want", V.SERVE_DNS)
def test_dnsbpf_serve_dns(zone, answers, want):
    types = {"A": dnsbpf.TYPE_A, "AAAA": dnsbpf.TYPE_AAAA}
    msg = dnsbpf.Msg([dnsbpf.RR(n, types[t], ttl, rd if t == "A" else "") for n, t, ttl, rd in answers])
    rec = RecordingMap()
    h = dnsbpf.Handler(zone, rec, next_handler=lambda req: msg)
    out = h.ServeDNS(object())
    assert out is msg and len(out.answer) == len(answers)                 # response forwarded unchanged
    assert rec.entries == [(ebpf.IPToUint32(ip), ebpf.DomainHash(hs), ttl) for ip, hs, ttl in want]
    assert dnsbpf.Handler(zone, None, next_handler=lambda req: msg).ServeDNS(object()) is msg   # nil Map: skip writes


def test_zone_to_domain():
    for z, d in V.ZONE_TO_DOMAIN:
        assert dnsbpf.zone_to_domain(z) == d
    assert dnsbpf.Handler("x.", None).Name() == "dnsbpf"


# ---- randomized: the C++ rule compiler (product) against the Python restatement ------------------------------
def _random_rules(rng, n):
    labels = ["api", "www", "cdn", "a", "b-c", "x_y", "GitHub", "example", "io", "com", "org", "co.uk", "internal", "docker"]
    protos = ["", "tls", "TLS", "tcp", "ssh", "http", "HTTP", "udp"]
    actions = ["", "allow", "Allow", "deny", "DENY", "log"]
    out = []
    for _ in range(n):
        u = rng.random()
        if u < 0.08:
            dst = "%d.%d.%d.%d" % tuple(rng.integers(1, 255, size=4))
        elif u < 0.16:
            dst = "%d.%d.%d.0/%d" % (*rng.integers(1, 255, size=3), int(rng.integers(8, 33)))
        elif u < 0.20:
            dst = ["", ".", "..", "*.example.com", "a b.com", "bad:port.com", "2001:db8::1", "::ffff:10.0.0.1", "10.0.0.0/33"][int(rng.integers(0, 9))]
        else:
            k = int(rng.integers(1, 4))
            dst = ".".join(labels[int(i)] for i in rng.integers(0, len(labels), size=k))
            if rng.random() < 0.25:
                dst = "." + dst
            if rng.random() < 0.1:
                dst += "."
        r = {"dst": dst, "proto": protos[int(rng.integers(0, len(protos)))], "port": int(rng.choice([0, 0, 22, 80, 443, 8080, 8443])),
             "action": actions[int(rng.integers(0, len(actions)))]}
        if rng.random() < 0.3:
            r["path_rules"] = [(["/", "/api", "/api/v1", "/admin", "", "/x?y"][int(rng.integers(0, 6))], actions[int(rng.integers(0, len(actions)))])
                               for _ in range(int(rng.integers(1, 4)))]
            r["path_default"] = ["", "deny", "allow", "ALLOW", "other"][int(rng.integers(0, 5))]
        out.append(r)
    return out


@pytest.mark.parametrize("seed", range(12))
def test_product_equals_oracle_on_random_rule_sets(eng, seed):
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    rules = _random_rules(rng, 120)
    rules += [dict(r) for r in rules[:10]]                                       # exact duplicates
    for r in rules:
        assert engine.validate_dst(r["dst"]) == ro.validate_dst(r["dst"]), r["dst"]
    c = firewall.compile_rules(eng, rules)
    o = ro.compile_rules(rules)
    assert c.zones == o.zones and c.server_names == o.snis and c.tcp_mappings == o.mappings and c.warnings == o.warnings
    got_http = []
    for line in eng.rules_dump(3):
        d, rt = line.split("|")
        got_http.append((d.split(","), [(x[1:], x[0] == "A") for x in rt.split(",")]))
    assert got_http == [(d, r) for d, r in o.http]
    assert sorted(eng.map_rows(7)) == sorted(int(a).to_bytes(4, "little") + int(m).to_bytes(4, "little") for a, m in set(o.cidrs))


def test_product_rule_normalisation_defaults(eng):
    """normalizeRule + ruleKey as the reference's rule store applies them before anything is generated
    (internal/firewall/rules.go:113-132; TestAddRules_DefaultProto / _NormalizesEmptyFields / _Deduplication /
    _DifferentPortsNotDuplicate, rules_test.go:62-102,236-258): an empty proto is tls, an empty action allow, a tls rule
    without a port is 443; explicit values are never overridden; the dedup key is dst:proto:port."""
    R = firewall.EgressRule
    c = firewall.compile_rules(eng, [R("a.com"), R("b.com", "ssh", 22)])
    assert "a.com" in c.server_names and "a.com" in c.zones                      # proto -> tls, action -> allow, port -> 443
    assert c.tcp_mappings == [("b.com", 22, 10001)] and "b.com" not in c.server_names      # ssh:22 stays what it is
    same = firewall.compile_rules(eng, [R("a.com"), R("a.com", "tls", 443, "allow"), R("a.com", "tls", 0, "")])
    assert same.server_names.count("a.com") == 1 and same.zones.count("a.com") == 1        # one rule after normalisation
    two = firewall.compile_rules(eng, [R("example.com", "tcp", 80, "allow"), R("example.com", "tcp", 443, "allow")])
    assert [(d, p) for d, p, _ in two.tcp_mappings] == [("example.com", 80), ("example.com", 443)]   # different ports are different rules
    assert len({ep for _, _, ep in two.tcp_mappings}) == 2
    # the oracle's restatement says the same
    o = ro.compile_rules([{"dst": "a.com"}, {"dst": "a.com", "proto": "tls", "port": 443, "action": "allow"}, {"dst": "b.com", "proto": "ssh", "port": 22}])
    assert list(o.snis).count("a.com") == 1 and [tuple(m)[:2] for m in o.mappings] == [("b.com", 22)]
