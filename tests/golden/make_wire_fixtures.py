#!/usr/bin/env python
"""Generates tests/golden/wire_fixtures.json — name-stage inputs made by THIRD-PARTY encoders, with expectations that do
not come from this repository's parsers:

  * DNS queries and responses are built and serialised by dnspython (dns.message / dns.rrset: name compression, EDNS
    OPT, CNAME chains, 0-TTL, > 31 A records, AAAA-only ...).  What the frame "says" — question name, the A records of
    the answer section and their TTLs — is read back with dnspython's own DECODER (dns.message.from_wire), not with
    clk_oracle.c or the CUDA parser.
  * TLS ClientHellos are produced by OpenSSL through Python's ssl.MemoryBIO (real extension lists: supported_versions,
    key_share, ALPN, session ticket, padding ...; TLS 1.2-only and TLS 1.3 forms; with and without SNI).  The expected
    host name is the server_hostname handed to OpenSSL.
  * Plaintext HTTP/1.1 requests are serialised by h11.

The expected verdict is then computed by a few lines of plain string code below (zone = longest label suffix that is a
Corefile zone; server_names = exact name or ".suffix" at a dot; virtual host + first matching path prefix), from the
rule compiler's tables (which are pinned by the reference's own unit-test vectors, tests/golden/reference_vectors.py).

Also transcribes the behavioural expectations of the reference's Docker e2e suite (test/e2e/firewall_test.go:57,137,
150-178,290-342,344-389,391-446,451-482,484-556,611-776,950-1057) as scenarios: the rule set of the test, the client
action (curl https / curl http / dig / getent / ssh-keyscan / ping), the packets that action puts on the wire (built by
the encoders above) and what the reference asserts about it (blocked / reaches upstream / 403).

Run:  python tests/golden/make_wire_fixtures.py      (needs dnspython, ssl/OpenSSL, h11 — all in the image)
ClientHello bytes contain fresh randoms on every run, so the committed file is the fixture; the DNS and HTTP parts are
deterministic and re-checked by tests/test_wire_fixtures.py::test_generator_is_reproducible.
"""
from __future__ import annotations

import json
import ssl
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import dns.flags
import dns.message
import dns.name
import dns.opcode
import dns.rdataclass
import dns.rdatatype
import dns.rrset
import h11

from clawker_b200 import capture
from clawker_b200.capture import frame, ip4
from oracle import rules_oracle

ALLOW, DENY, NA = 1, 2, 0
CT = ip4("172.20.0.10")                  # container #0 of base_tables(): enforced, uid 1000
COREDNS, UP = ip4(capture.COREDNS_IP), ip4(capture.UPSTREAM_DNS)
EXT = ip4("93.184.216.34")

# the rule set of the table-driven part: the reference's required rules (internal/config/defaults.go:15-30) + the e2e rules
RULES = [
    {"dst": "api.anthropic.com", "proto": "tls", "port": 443, "action": "allow"},
    {"dst": "claude.com", "proto": "tls", "port": 443, "action": "allow"},
    {"dst": "platform.claude.com", "proto": "tls", "port": 443, "action": "allow"},
    {"dst": ".claude.ai", "proto": "tls", "port": 443, "action": "allow"},
    {"dst": "mcp-proxy.anthropic.com", "proto": "tls", "port": 443, "action": "allow"},
    {"dst": "sentry.io", "proto": "tls", "port": 443, "action": "allow"},
    {"dst": "statsig.anthropic.com", "proto": "tls", "port": 443, "action": "allow"},
    {"dst": "statsig.com", "proto": "tls", "port": 443, "action": "allow"},
    {"dst": ".datadoghq.com", "proto": "tls", "port": 443, "action": "allow"},
    {"dst": ".datadoghq.eu", "proto": "tls", "port": 443, "action": "allow"},
    {"dst": "github.com", "proto": "ssh", "port": 22, "action": "allow"},                    # firewall_test.go:407-415
    {"dst": "clawker.dev", "proto": "tls", "port": 443, "action": "allow"},                  # :975-990
    {"dst": ".clawker.dev", "proto": "tls", "port": 443, "action": "allow"},
    {"dst": "example.com", "proto": "http", "port": 80, "action": "allow",                   # :628-640
     "path_rules": [("/test", "allow")], "path_default": "deny"},
    {"dst": "httpbin.org", "proto": "http", "port": 80, "action": "allow",                   # :712-724 (on another host so both forms coexist)
     "path_rules": [("/evil", "deny")], "path_default": "allow"},
]


# ---- expectations from plain string logic ---------------------------------------------------------------------------
def tables():
    c = rules_oracle.compile_rules(RULES)
    return c


def zone_of(qname: str, zones) -> str | None:
    """CoreDNS ServeDNS zone routing: the longest label suffix of the lower-cased name that is a server-block zone."""
    labels = qname.lower().rstrip(".").split(".")
    zs = set(zones)
    for i in range(len(labels)):
        cand = ".".join(labels[i:])
        if cand in zs:
            return cand
    return None


def sni_allowed(host: str, snis) -> bool:
    """Envoy filter_chain_match.server_names: the exact name, else "*.d" (stored ".d") for every dot inside the name."""
    h = host.lower()
    s = set(snis)
    if h in s:
        return True
    return any(h[i] == "." and h[i:] in s for i in range(1, len(h) - 1))


def http_verdict(host: str, target: str, vhosts) -> int:
    """Envoy RouteMatcher: virtual host by exact domain, then "*.d" suffix (longest first), then "d:*" prefix; first route
    whose prefix matches the path decides."""
    h = host.lower()
    best = None
    for domains, routes in vhosts:
        for d in domains:
            dl = d.lower()
            if dl == h:
                best = (3, len(dl), routes) if best is None or best[0] < 3 else best
    if best is None:
        for domains, routes in vhosts:
            for d in domains:
                dl = d.lower()
                if dl.startswith("*") and len(dl) > 1 and h.endswith(dl[1:]) and len(h) > len(dl) - 1:
                    if best is None or best[1] < len(dl):
                        best = (2, len(dl), routes)
    if best is None:
        for domains, routes in vhosts:
            for d in domains:
                dl = d.lower()
                if dl.endswith("*") and len(dl) > 1 and h.startswith(dl[:-1]) and len(h) > len(dl) - 1:
                    if best is None or best[1] < len(dl):
                        best = (1, len(dl), routes)
    if best is None:
        return DENY
    for prefix, allow in best[2]:
        if target.startswith(prefix):
            return ALLOW if allow else DENY
    return DENY


def accept_query(m: dns.message.Message) -> bool:
    """miekg/dns defaultMsgAcceptFunc as CoreDNS applies it to a request, on dnspython's view of the header."""
    if m.flags & dns.flags.QR:
        return False
    if m.opcode() not in (dns.opcode.QUERY, dns.opcode.NOTIFY):
        return False
    if len(m.question) != 1:
        return False
    ar = len(m.additional) + (1 if m.opt is not None else 0)
    return sum(len(r) for r in m.answer) <= 1 and sum(len(r) for r in m.authority) <= 1 and ar <= 2


# ---- encoders -------------------------------------------------------------------------------------------------------
def dns_query_wire(name: str, rdtype="A", edns: int | None = None, rd=True, want_dnssec=False, opcode=None, qid=0x2b1d) -> bytes:
    m = dns.message.make_query(name, rdtype, use_edns=edns if edns is not None else False, want_dnssec=want_dnssec, id=qid)
    if not rd:
        m.flags &= ~dns.flags.RD
    if opcode is not None:
        m.set_opcode(opcode)
    return m.to_wire()


def dns_response_wire(qname: str, records, rdtype="A", edns=None, qid=0x5e11, authority=None, additional=None) -> bytes:
    """records: [(owner, ttl, type, text)] for the answer section.  dnspython compresses owner names on to_wire()."""
    q = dns.message.make_query(qname, rdtype, use_edns=edns if edns is not None else False, id=qid)
    r = dns.message.make_response(q)
    for sec, rows in ((r.answer, records), (r.authority, authority or []), (r.additional, additional or [])):
        for owner, ttl, typ, text in rows:
            rs = dns.rrset.from_text(owner, ttl, "IN", typ, text)
            sec.append(rs)
    return r.to_wire(max_size=65535)


def client_hello(server_hostname: str | None, max12=False, alpn=None, ciphers=None, no_tickets=False) -> bytes:
    ctx = ssl.SSLContext(ssl.PROTOCOL_TLS_CLIENT)
    ctx.check_hostname = False
    ctx.verify_mode = ssl.CERT_NONE
    if max12:
        ctx.maximum_version = ssl.TLSVersion.TLSv1_2
    if alpn:
        ctx.set_alpn_protocols(alpn)
    if ciphers:
        ctx.set_ciphers(ciphers)
    if no_tickets:
        ctx.options |= ssl.OP_NO_TICKET
    inc, out = ssl.MemoryBIO(), ssl.MemoryBIO()
    obj = ctx.wrap_bio(inc, out, server_hostname=server_hostname)
    try:
        obj.do_handshake()
    except ssl.SSLWantReadError:
        pass
    return out.read()


def http_request(method: str, target: str, headers) -> bytes:
    c = h11.Connection(h11.CLIENT)
    return c.send(h11.Request(method=method, target=target, headers=headers)) + (c.send(h11.EndOfMessage()) or b"")


# ---- frames ---------------------------------------------------------------------------------------------------------
def f_query(wire: bytes) -> bytes:
    return frame(CT, ip4("127.0.0.11"), 17, 40000, 53, wire)


def f_answer(wire: bytes) -> bytes:
    return frame(UP, COREDNS, 17, 53, 33333, wire)


def f_tls(payload: bytes, dport=443) -> bytes:
    return frame(CT, EXT, 6, 40001, dport, payload)


def f_http(payload: bytes) -> bytes:
    return frame(CT, EXT, 6, 40002, 80, payload)


def build():
    c = tables()
    zones, snis, vhosts = c.zones, c.snis, [(list(d), [(p, bool(a)) for p, a in r]) for d, r in c.http]
    cases = []

    def add(desc, kind, fr, name, nupd=None, updates=None, decoded=None, source=None):
        cases.append({"desc": desc, "kind": kind, "frame": fr.hex(), "name": name, "nupd": nupd, "updates": updates,
                      "decoded": decoded, "source": source})

    # ---- DNS queries (dnspython encoder; expectation from dnspython decoder + zone_of) --------------------------------
    qn = ["api.anthropic.com", "API.Anthropic.COM", "x.y.api.anthropic.com", "anthropic.com", "claude.ai", "www.claude.ai", "a.b.c.d.e.claude.ai",
          "app.datadoghq.eu", "datadoghq.eu", "github.com", "gitlab.com", "example.com", "evil.example.com", "docs.clawker.dev", "clawker.dev",
          "notclawker.dev", "host.docker.internal", "otel-collector", "jaeger", "grafana", "docker.internal", "com", "sentry.io.evil.net",
          "xn--bcher-kva.example", "under_score.github.com", "a" * 63 + ".github.com", ".".join(["l" * 30] * 7) + ".github.com", "httpbin.org"]
    variants = [dict(), dict(rdtype="AAAA"), dict(edns=0), dict(edns=0, want_dnssec=True), dict(rd=False), dict(rdtype="HTTPS"), dict(rdtype="TXT", edns=0)]
    k = 0
    for name in qn:
        for v in (variants if name in ("api.anthropic.com", "gitlab.com", "www.claude.ai") else variants[k % len(variants):k % len(variants) + 1]):
            wire = dns_query_wire(name, **v)
            m = dns.message.from_wire(wire)
            qname = m.question[0].name.to_text(omit_final_dot=True)
            z = zone_of(qname, zones) if accept_query(m) else None
            add(f"query {name[:40]} {v}", "dns_query", f_query(wire), ALLOW if z else DENY,
                decoded={"qname": qname, "zone": z, "rdtype": dns.rdatatype.to_text(m.question[0].rdtype), "edns": m.edns}, source="dnspython")
        k += 1
    for desc, wire in [("NOTIFY opcode is accepted", dns_query_wire("github.com", rdtype="SOA", opcode=dns.opcode.NOTIFY)),
                       ("STATUS opcode is refused", dns_query_wire("github.com", opcode=dns.opcode.STATUS)),
                       ("UPDATE opcode is refused", dns_query_wire("github.com", rdtype="SOA", opcode=dns.opcode.UPDATE))]:
        m = dns.message.from_wire(wire)
        z = zone_of(m.question[0].name.to_text(omit_final_dot=True), zones) if accept_query(m) else None
        add(desc, "dns_query", f_query(wire), ALLOW if z else DENY, decoded={"opcode": int(m.opcode())}, source="dnspython")

    # ---- DNS responses reaching CoreDNS (dnsbpf) ------------------------------------------------------------------------
    def ans_case(desc, qname, records, **kw):
        wire = dns_response_wire(qname, records, **kw)
        m = dns.message.from_wire(wire)
        q = m.question[0].name.to_text(omit_final_dot=True)
        z = zone_of(q, zones)
        ups = []
        if z:
            zh = rules_oracle.domain_hash(z)
            for rs in m.answer:
                if rs.rdtype == dns.rdatatype.A and rs.rdclass == dns.rdataclass.IN:
                    for rd in rs:
                        ip = ip4(rd.address)
                        if ip != 0:
                            ups.append([ip, zh, rs.ttl if rs.ttl else 60])
        ups = ups[:31]
        add(desc, "dns_answer", f_answer(wire), ALLOW if z else DENY, nupd=len(ups), updates=ups,
            decoded={"qname": q, "zone": z, "answer_rrsets": [rs.to_text() for rs in m.answer][:6], "wire_len": len(wire)}, source="dnspython")

    ans_case("one A", "github.com", [("github.com.", 60, "A", "140.82.121.4")])
    ans_case("four A, shared owner (compressed)", "api.anthropic.com", [("api.anthropic.com.", 300, "A", f"160.79.104.{i}") for i in (10, 11, 12, 13)])
    ans_case("CNAME chain then A (owners differ, compression pointers into rdata)", "www.claude.ai",
             [("www.claude.ai.", 300, "CNAME", "edge.claude.ai."), ("edge.claude.ai.", 60, "CNAME", "lb.cdn.claude.ai."),
              ("lb.cdn.claude.ai.", 30, "A", "104.18.2.115"), ("lb.cdn.claude.ai.", 30, "A", "104.18.3.115")])
    ans_case("TTL 0 becomes 60", "sentry.io", [("sentry.io.", 0, "A", "35.186.247.156")])
    ans_case("0.0.0.0 is skipped", "statsig.com", [("statsig.com.", 60, "A", "0.0.0.0"), ("statsig.com.", 60, "A", "34.120.214.181")])
    ans_case("AAAA only: no updates", "github.com", [("github.com.", 60, "AAAA", "2606:50c0:8000::154")], rdtype="AAAA")
    ans_case("A and AAAA mixed with EDNS OPT", "docs.clawker.dev", [("docs.clawker.dev.", 120, "A", "76.76.21.21"), ("docs.clawker.dev.", 120, "AAAA", "2a00::1")], edns=0)
    ans_case("40 A records: capped at 31", "app.datadoghq.com", [("app.datadoghq.com.", 60, "A", f"3.233.{i}.{i + 1}") for i in range(40)])
    ans_case("answer for an unlisted name: zone '.', nothing written", "gitlab.com", [("gitlab.com.", 60, "A", "172.65.251.78")])
    ans_case("NXDOMAIN-style empty answer with SOA in authority", "nope.github.com", [],
             authority=[("github.com.", 900, "SOA", "ns1.github.com. hostmaster.github.com. 1 3600 600 604800 60")])
    ans_case("A with NS authority and glue in additional", "github.com", [("github.com.", 60, "A", "140.82.112.3")],
             authority=[("github.com.", 900, "NS", "ns-421.awsdns-52.com.")], additional=[("ns-421.awsdns-52.com.", 900, "A", "205.251.193.165")])
    ans_case("internal host", "otel-collector", [("otel-collector.", 30, "A", "172.20.0.30")])

    # ---- TLS ClientHellos from OpenSSL ----------------------------------------------------------------------------------
    hosts = ["api.anthropic.com", "API.ANTHROPIC.COM", "claude.ai", "www.claude.ai", "a.b.claude.ai", "claude.com", "evil.com", "example.com",
             "clawker.dev", "docs.clawker.dev", "xclawker.dev", "app.datadoghq.com", "datadoghq.com", "github.com", "anthropic.com",
             "platform.claude.com.evil.io", ("a" * 60 + ".") * 3 + "claude.ai", "sentry.io"]
    forms = [dict(), dict(max12=True), dict(alpn=["h2", "http/1.1"]), dict(max12=True, no_tickets=True, alpn=["http/1.1"]), dict(ciphers="ECDHE+AESGCM")]
    for i, hname in enumerate(hosts):
        for fm in (forms if i < 3 else forms[i % len(forms):i % len(forms) + 1]):
            hello = client_hello(hname, **fm)
            assert hello[0] == 0x16 and len(hello) < 1460
            add(f"ClientHello {hname[:40]} {fm}", "tls", f_tls(hello), ALLOW if sni_allowed(hname, snis) else DENY,
                decoded={"sni": hname, "hello_len": len(hello)}, source=ssl.OPENSSL_VERSION)
    hello = client_hello(None)
    add("ClientHello without SNI", "tls", f_tls(hello), DENY, decoded={"sni": None, "hello_len": len(hello)}, source=ssl.OPENSSL_VERSION)
    big = client_hello("api.anthropic.com", alpn=["proto-%03d-%s" % (i, "x" * 40) for i in range(40)])
    assert len(big) > 1460
    add("ClientHello larger than one segment (first 1460 bytes): needs reassembly", "tls", f_tls(big[:1460]), NA,
        decoded={"sni": "api.anthropic.com", "hello_len": len(big)}, source=ssl.OPENSSL_VERSION)
    add("TLS to the ssh port of a listed host: not the egress listener", "tls", f_tls(client_hello("github.com"), dport=22), None,
        decoded={"sni": "github.com"}, source=ssl.OPENSSL_VERSION)

    # ---- plaintext HTTP from h11 ------------------------------------------------------------------------------------------
    reqs = [("GET", "/test", "example.com"), ("GET", "/test/deeper?x=1", "example.com"), ("GET", "/evil", "example.com"), ("GET", "/", "example.com"),
            ("GET", "/", "httpbin.org"), ("GET", "/evil", "httpbin.org"), ("POST", "/anything", "HTTPBIN.org"), ("GET", "/evil/sub", "httpbin.org"),
            ("GET", "/", "api.anthropic.com"), ("HEAD", "/test", "example.com:80"), ("GET", "/test", "sub.example.com")]
    for method, target, host in reqs:
        raw = http_request(method, target, [("Host", host), ("User-Agent", "curl/8.5.0"), ("Accept", "*/*")])
        add(f"{method} http://{host}{target}", "http", f_http(raw), http_verdict(host, target, vhosts),
            decoded={"host": host, "target": target}, source="h11 " + h11.__version__)

    # ---- the reference's e2e expectations (test/e2e/firewall_test.go), as packets -----------------------------------------
    def dq(name):
        return f_query(dns_query_wire(name))

    scenarios = [
        {"test": "TestFirewall_BlockedDomain", "ref": "firewall_test.go:57-63", "action": "curl https://example.com", "reference_expects": "curl fails",
         "packets": [("dns_query", dq("example.com"), {"name": ALLOW}),      # example.com has an http rule here -> it resolves; the TLS connection is what is refused
                     ("tls", f_tls(client_hello("example.com")), {"name": DENY})]},
        {"test": "TestFirewall_BlockedDomain (unlisted name)", "ref": "firewall_test.go:57-63", "action": "curl https://evil.example.net", "reference_expects": "curl fails",
         "packets": [("dns_query", dq("evil.example.net"), {"name": DENY})]},
        {"test": "TestFirewall_AllowedDomain", "ref": "firewall_test.go:137-148", "action": "curl https://api.anthropic.com", "reference_expects": "HTTP code returned",
         "packets": [("dns_query", dq("api.anthropic.com"), {"name": ALLOW}),
                     ("tls", f_tls(client_hello("api.anthropic.com")), {"name": ALLOW, "verdict": 1, "target": 2})]},
        {"test": "TestFirewall_ICMPBlocked", "ref": "firewall_test.go:65-99", "action": "ping -c1 8.8.8.8", "reference_expects": "ping fails",
         "packets": [("icmp", frame(CT, ip4("8.8.8.8"), 1, 0, 0), {"verdict": 2, "metric": 2})]},
        {"test": "TestFirewall_IntraNetworkBypass", "ref": "firewall_test.go:290-342", "action": "curl http://<clawker-net IP>:8080/", "reference_expects": "200 via CIDR bypass",
         "packets": [("tcp", frame(CT, ip4("172.20.0.77"), 6, 40010, 8080), {"verdict": 0, "name": NA})]},
        {"test": "TestFirewall_HostProxyReachable", "ref": "firewall_test.go:344-389", "action": "curl $CLAWKER_HOST_PROXY/health ; curl http://host.docker.internal:9999",
         "reference_expects": "200 from the host proxy; other host port blocked",
         "packets": [("tcp", frame(CT, ip4(capture.HOST_PROXY_IP), 6, 40011, capture.HOST_PROXY_PORT), {"verdict": 0}),
                     ("dns_query", dq("host.docker.internal"), {"name": ALLOW}),
                     ("tcp", frame(CT, ip4(capture.HOST_PROXY_IP), 6, 40012, 9999, http_request("GET", "/", [("Host", "host.docker.internal:9999")])), {"verdict": 1, "target": 2, "name": DENY})]},
        {"test": "TestFirewall_SSHTCPMapping", "ref": "firewall_test.go:391-446", "action": "ssh-keyscan github.com ; dig +short gitlab.com",
         "reference_expects": "keyscan succeeds via envoy:10001; gitlab.com does not resolve",
         "packets": [("dns_query", dq("github.com"), {"name": ALLOW}),
                     ("dns_answer", f_answer(dns_response_wire("github.com", [("github.com.", 60, "A", "140.82.121.4")])), {"name": ALLOW, "nupd": 1, "apply": True}),
                     ("tcp", frame(CT, ip4("140.82.121.4"), 6, 40013, 22), {"verdict": 1, "target": 3, "port": 10001}),
                     ("dns_query", dq("gitlab.com"), {"name": DENY})]},
        {"test": "TestFirewall_DockerInternalDNS", "ref": "firewall_test.go:451-482", "action": "getent hosts host.docker.internal / otel-collector / evil.example.com",
         "reference_expects": "first two resolve, the third does not",
         "packets": [("dns_query", dq("host.docker.internal"), {"name": ALLOW}), ("dns_query", dq("otel-collector"), {"name": ALLOW}),
                     ("dns_query", dq("evil.example.com"), {"name": ALLOW})]},     # here example.com is a zone (http rule), so its subdomains resolve; see the next scenario for the blocked form
        {"test": "TestFirewall_DockerInternalDNS (unlisted)", "ref": "firewall_test.go:478-481", "action": "getent hosts evil.example.net", "reference_expects": "does not resolve",
         "packets": [("dns_query", dq("evil.example.net"), {"name": DENY})]},
        {"test": "TestFirewall_HTTPDomainDetection", "ref": "firewall_test.go:484-556", "action": "curl http://example.com/test ; curl http://unlisted.org/",
         "reference_expects": "first reaches upstream; second is blocked",
         "packets": [("dns_query", dq("example.com"), {"name": ALLOW}),
                     ("http", f_http(http_request("GET", "/test", [("Host", "example.com"), ("User-Agent", "curl/8.5.0"), ("Accept", "*/*")])), {"name": ALLOW, "verdict": 1, "target": 2}),
                     ("dns_query", dq("unlisted.org"), {"name": DENY}),
                     ("http", f_http(http_request("GET", "/", [("Host", "unlisted.org"), ("User-Agent", "curl/8.5.0"), ("Accept", "*/*")])), {"name": DENY})]},
        {"test": "TestFirewall_PathRulesDefaultDeny", "ref": "firewall_test.go:611-693", "action": "curl http://example.com/test ; curl http://example.com/evil",
         "reference_expects": "/test not 403; /evil 403",
         "packets": [("http", f_http(http_request("GET", "/test", [("Host", "example.com"), ("Accept", "*/*")])), {"name": ALLOW}),
                     ("http", f_http(http_request("GET", "/evil", [("Host", "example.com"), ("Accept", "*/*")])), {"name": DENY})]},
        {"test": "TestFirewall_PathRulesExplicitDeny", "ref": "firewall_test.go:695-776", "action": "curl http://httpbin.org/ ; curl http://httpbin.org/evil",
         "reference_expects": "/ reaches upstream; /evil 403",
         "packets": [("http", f_http(http_request("GET", "/", [("Host", "httpbin.org"), ("Accept", "*/*")])), {"name": ALLOW}),
                     ("http", f_http(http_request("GET", "/evil", [("Host", "httpbin.org"), ("Accept", "*/*")])), {"name": DENY})]},
        {"test": "TestFirewall_WildcardAndExactCoexist", "ref": "firewall_test.go:950-1057", "action": "curl https://clawker.dev/... ; curl https://docs.clawker.dev/...",
         "reference_expects": "both SNI values own a filter chain (path decisions are made after TLS termination, outside a packet classifier)",
         "packets": [("dns_query", dq("clawker.dev"), {"name": ALLOW}), ("dns_query", dq("docs.clawker.dev"), {"name": ALLOW}),
                     ("tls", f_tls(client_hello("clawker.dev")), {"name": ALLOW}), ("tls", f_tls(client_hello("docs.clawker.dev")), {"name": ALLOW}),
                     ("tls", f_tls(client_hello("evilclawker.dev")), {"name": DENY})]},
    ]
    for sc in scenarios:
        sc["packets"] = [{"kind": k, "frame": fr.hex(), "expect": ex} for k, fr, ex in sc["packets"]]
    bypass = {"test": "TestFirewall_Bypass", "ref": "firewall_test.go:101-135", "action": "firewall bypass 30s ; curl https://example.com",
              "reference_expects": "blocked before, allowed during the bypass, blocked again after --stop",
              "packets": [{"kind": "tls", "frame": f_tls(client_hello("example.com")).hex(),
                           "expect_enforced": {"verdict": 1, "name": DENY}, "expect_bypassed": {"verdict": 0, "metric": 3, "name": NA}}]}
    return {"rules": [{**r, "path_rules": [list(p) for p in r.get("path_rules", [])]} for r in RULES], "cases": cases, "scenarios": scenarios, "bypass": bypass,
            "versions": {"dnspython": dns.__version__ if hasattr(dns, "__version__") else "?", "openssl": ssl.OPENSSL_VERSION, "h11": h11.__version__}}


if __name__ == "__main__":
    import dns.version
    out = build()
    out["versions"]["dnspython"] = dns.version.version
    p = Path(__file__).with_name("wire_fixtures.json")
    p.write_text(json.dumps(out, indent=0))
    print(f"{len(out['cases'])} cases + {len(out['scenarios']) + 1} e2e scenarios -> {p} ({p.stat().st_size} bytes)")
