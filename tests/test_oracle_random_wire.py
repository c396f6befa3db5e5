"""Randomised pinning of the oracle's name stages (CPU only): thousands of seeded random messages from THIRD-PARTY encoders — dnspython
queries / responses (compression, EDNS, CNAME chains, AAAA mixed in), OpenSSL ClientHellos (ssl.MemoryBIO), h11 requests — with the
expectation from the third-party decoder plus plain string matching (the helpers of tests/golden/make_wire_fixtures.py).  The committed
fixture (tests/test_wire_fixtures.py) holds the hand-picked cases and is also replayed on the GPU; this file widens the oracle's side."""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import pytest

from clawker_b200 import abi, capture
from oracle import oracle, rules_oracle

dns = pytest.importorskip("dns.message")
import dns.message  # noqa: E402
import dns.rdataclass  # noqa: E402
import dns.rdatatype  # noqa: E402

_spec = importlib.util.spec_from_file_location("make_wire_fixtures", Path(__file__).resolve().parent / "golden" / "make_wire_fixtures.py")
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)
C = G.tables()
ALPHA = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789-_"


def _tables():
    t = capture.base_tables(8)
    t.update(zones=C.zones, snis=C.snis, routes=dict(C.routes), cidrs=list(C.cidrs), http=list(C.http))
    return t


def _rname(rng, pool, hostlike=False):
    alpha = ALPHA[:-1] if hostlike else ALPHA                       # '_' is not a host-name character for OpenSSL / h11 users
    labs = ["".join(alpha[int(i)] for i in rng.integers(0, len(alpha), size=int(rng.integers(1, 20)))).strip("-") or "x" for _ in range(int(rng.integers(1, 6)))]
    base = ".".join(labs)
    if rng.random() < 0.5:
        z = pool[int(rng.integers(0, len(pool)))].lstrip(".")
        z = "".join(ch.upper() if rng.random() < 0.3 else ch for ch in z)
        return (base + "." + z) if rng.random() < 0.7 else z
    return base


def _classify(frames):
    blob, offs = capture.build_varlen(frames)
    o = oracle.Oracle("port").load(_tables())
    w, _ = o.classify(blob, offs, flags=abi.F_NAME_STAGE)
    return abi.verdict_fields(w), o.dns_answers(blob, offs)


def test_random_dns_messages_against_dnspython():
    rng = np.random.Generator(np.random.PCG64(77))
    zl = sorted(C.zones)
    frames, exp = [], []
    for _ in range(1500):
        name = _rname(rng, zl)
        v = [dict(), dict(rdtype="AAAA"), dict(edns=0), dict(rd=False), dict(rdtype="TXT", edns=0, want_dnssec=True)][int(rng.integers(0, 5))]
        wire = G.dns_query_wire(name, **v)
        m = dns.message.from_wire(wire)
        z = G.zone_of(m.question[0].name.to_text(omit_final_dot=True), C.zones) if G.accept_query(m) else None
        frames.append(G.f_query(wire)); exp.append((G.ALLOW if z else G.DENY, None))
    for _ in range(800):
        name = _rname(rng, zl)
        recs, owner = [], name + "."
        if rng.random() < 0.3:
            tgt = "edge-" + _rname(rng, zl) + "."
            recs.append((owner, int(rng.integers(0, 600)), "CNAME", tgt)); owner = tgt
        attl = int(rng.choice([0, 30, 60, 300, 86400]))            # one TTL per RRset (RFC 2181 5.2; dnspython merges an owner's RRs into one rrset)
        for _k in range(int(rng.integers(0, 6))):
            if rng.random() < 0.75:
                recs.append((owner, attl, "A", ".".join(str(int(x)) for x in rng.integers(1, 255, size=4))))
            else:
                recs.append((owner, 60, "AAAA", "2606:4700::%x" % int(rng.integers(1, 65535))))
        wire = G.dns_response_wire(name, recs, edns=0 if rng.random() < 0.3 else None)
        m = dns.message.from_wire(wire)
        z = G.zone_of(m.question[0].name.to_text(omit_final_dot=True), C.zones)
        ups = []
        if z:
            zh = rules_oracle.domain_hash(z)
            for rs in m.answer:
                if rs.rdtype == dns.rdatatype.A and rs.rdclass == dns.rdataclass.IN:
                    ups += [[G.ip4(rd.address), zh, rs.ttl if rs.ttl else 60] for rd in rs if G.ip4(rd.address) != 0]
        frames.append(G.f_answer(wire)); exp.append((G.ALLOW if z else G.DENY, ups[:31]))
    f, upd = _classify(frames)
    bad = []
    for i, (nm, ups) in enumerate(exp):
        got = [[int(r["ip"]), int(r["domain_hash"]), int(r["ttl"])] for r in upd[upd["packet"] == i]] if ups is not None else None
        if int(f["name"][i]) != nm or (ups is not None and (got != ups or int(f["nupd"][i]) != len(ups))):
            bad.append((i, int(f["name"][i]), nm, got, ups))
    assert not bad, bad[:5]
    assert sum(1 for nm, _ in exp if nm == G.ALLOW) > 500 and sum(1 for nm, _ in exp if nm == G.DENY) > 500


def test_random_client_hellos_against_openssl():
    rng = np.random.Generator(np.random.PCG64(78))
    pool = sorted(C.snis)
    frames, exp = [], []
    for k in range(250):
        host = _rname(rng, pool, hostlike=True)[:200]
        kw = [dict(), dict(max12=True), dict(alpn=["h2", "http/1.1"]), dict(no_tickets=True), dict(alpn=["h2"], max12=True)][k % 5]
        try:
            hello = G.client_hello(host, **kw)
        except Exception:                                            # a name OpenSSL refuses (label too long ...) is not a case
            continue
        if len(hello) >= 1460:
            continue
        frames.append(G.f_tls(hello)); exp.append(G.ALLOW if G.sni_allowed(host, C.snis) else G.DENY)
    f, _ = _classify(frames)
    bad = [(i, int(f["name"][i]), e) for i, e in enumerate(exp) if int(f["name"][i]) != e]
    assert len(exp) > 200 and not bad, bad[:5]
    assert exp.count(G.ALLOW) > 40 and exp.count(G.DENY) > 40


def test_random_http_requests_against_h11():
    h11 = pytest.importorskip("h11")
    rng = np.random.Generator(np.random.PCG64(79))
    vhosts = [(list(d), [(p, bool(a)) for p, a in r]) for d, r in C.http]
    doms = sorted({x.lstrip("*").lstrip(".").rstrip("*").rstrip(":") for d, _ in vhosts for x in d if x not in ("*",)})
    frames, exp = [], []
    paths = ["/", "/test", "/test/a/b", "/tes", "/evil", "/evil/x", "/Test", "/anything?q=1", "/api/v1"]
    for _ in range(400):
        host = _rname(rng, doms, hostlike=True)[:120]
        if rng.random() < 0.2:
            host += ":%d" % int(rng.choice([80, 8080]))
        target = paths[int(rng.integers(0, len(paths)))]
        hdrs = [("User-Agent", "curl/8.5.0"), ("Accept", "*/*")][: int(rng.integers(0, 3))]
        pos = int(rng.integers(0, len(hdrs) + 1))
        hdrs = hdrs[:pos] + [("hOsT" if rng.random() < 0.3 else "Host", host)] + hdrs[pos:]
        try:
            wire = G.http_request(str(rng.choice(["GET", "POST", "HEAD"])), target, hdrs)
        except h11.LocalProtocolError:
            continue
        frames.append(G.f_http(wire)); exp.append(G.http_verdict(host, target, vhosts))
    f, _ = _classify(frames)
    bad = [(i, int(f["name"][i]), e, frames[i][54:120]) for i, e in enumerate(exp) if int(f["name"][i]) != e]
    assert len(exp) > 300 and not bad, bad[:5]
    assert exp.count(G.ALLOW) > 20 and exp.count(G.DENY) > 100
