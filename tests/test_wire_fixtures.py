"""Name stages against inputs made by THIRD-PARTY encoders (dnspython, OpenSSL via ssl.MemoryBIO, h11) with expectations
from dnspython's decoder / the encoders' own inputs and plain string matching — see tests/golden/make_wire_fixtures.py —
plus the behavioural expectations of the reference's Docker e2e suite (test/e2e/firewall_test.go) replayed as packets.
The committed fixture is checked against the CPU oracle here and against the CUDA path on the B200 (-m gpu)."""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np
import pytest

from clawker_b200 import abi, capture
from oracle import oracle, rules_oracle

GOLD = Path(__file__).resolve().parent / "golden"
FX = json.loads((GOLD / "wire_fixtures.json").read_text())
RULES = [{**r, "path_rules": [tuple(p) for p in r.get("path_rules", [])]} for r in FX["rules"]]
F = abi.F_NAME_STAGE


def _oracle_tables(extra_dns=None):
    t = capture.base_tables(8)
    c = rules_oracle.compile_rules(RULES)
    t.update(zones=c.zones, snis=c.snis, routes=dict(c.routes), cidrs=list(c.cidrs), http=list(c.http))
    if extra_dns:
        t["dns_cache"] = {**t["dns_cache"], **extra_dns}
    return t


def _frames(items):
    return capture.build_varlen([bytes.fromhex(x["frame"]) for x in items])


def _check_cases(words, upd):
    f = abi.verdict_fields(words)
    bad = []
    for i, c in enumerate(FX["cases"]):
        if c["name"] is not None and int(f["name"][i]) != c["name"]:
            bad.append(f"{c['desc']} [{c['source']}]: name verdict {int(f['name'][i])}, expected {c['name']} (decoded {c['decoded']})")
        if c["nupd"] is not None:
            got = [[int(r["ip"]), int(r["domain_hash"]), int(r["ttl"])] for r in upd[upd["packet"] == i]]
            if int(f["nupd"][i]) != c["nupd"] or got != c["updates"]:
                bad.append(f"{c['desc']}: updates {got} (word says {int(f['nupd'][i])}), dnspython says {c['updates']}")
    assert not bad, "\n".join(bad[:12]) + f"\n... {len(bad)} of {len(FX['cases'])} cases differ"


def test_fixture_covers_every_source():
    kinds = {c["kind"] for c in FX["cases"]}
    assert kinds == {"dns_query", "dns_answer", "tls", "http"} and len(FX["cases"]) >= 100
    assert any("OpenSSL" in (c["source"] or "") for c in FX["cases"]) and any(c["source"] == "dnspython" for c in FX["cases"])
    assert sum(1 for c in FX["cases"] if c["kind"] == "dns_answer" and c["nupd"] == 31) == 1
    assert {"TestFirewall_BlockedDomain", "TestFirewall_AllowedDomain", "TestFirewall_SSHTCPMapping", "TestFirewall_HTTPDomainDetection",
            "TestFirewall_PathRulesDefaultDeny", "TestFirewall_PathRulesExplicitDeny", "TestFirewall_WildcardAndExactCoexist",
            "TestFirewall_DockerInternalDNS", "TestFirewall_ICMPBlocked", "TestFirewall_IntraNetworkBypass", "TestFirewall_HostProxyReachable"} <= {s["test"] for s in FX["scenarios"]}


def test_generator_is_reproducible():
    """The dnspython and h11 parts of the fixture are deterministic: regenerate them and compare byte for byte (ClientHellos
    carry fresh randoms, so only their count and expectation are compared)."""
    sys.path.insert(0, str(GOLD))
    import make_wire_fixtures as G
    fresh = G.build()
    assert len(fresh["cases"]) == len(FX["cases"]) and fresh["rules"] == FX["rules"]
    for a, b in zip(fresh["cases"], FX["cases"]):
        assert (a["desc"], a["kind"], a["name"], a["nupd"], a["updates"]) == (b["desc"], b["kind"], b["name"], b["nupd"], b["updates"])
        if a["kind"] != "tls":
            assert a["frame"] == b["frame"], a["desc"]


def test_oracle_agrees_with_third_party_decoders():
    blob, offs = _frames(FX["cases"])
    o = oracle.Oracle("port").load(_oracle_tables())
    v, _ = o.classify(blob, offs, flags=F)
    _check_cases(v, o.dns_answers(blob, offs))


def _check_scenarios(classify, apply_updates):
    """classify(blob, offs) -> words; apply_updates(blob, offs, words) writes the batch's dns_cache updates."""
    bad = []
    for sc in FX["scenarios"]:
        for k, p in enumerate(sc["packets"]):
            blob, offs = _frames([p])
            w = classify(blob, offs)
            f = abi.verdict_fields(w)
            ex = dict(p["expect"])
            if ex.pop("apply", False):
                apply_updates(blob, offs, w)
            for field, want in ex.items():
                if int(f[field][0]) != want:
                    bad.append(f"{sc['test']} ({sc['ref']}) packet {k} [{p['kind']}]: {field} = {int(f[field][0])}, expected {want}  -- {sc['action']}: {sc['reference_expects']}")
    assert not bad, "\n".join(bad)


def test_oracle_replays_the_reference_e2e_expectations():
    state = {"extra": {}}

    def classify(blob, offs):
        return oracle.Oracle("port").load(_oracle_tables(state["extra"])).classify(blob, offs, flags=F)[0]

    def apply_updates(blob, offs, w):
        for r in oracle.Oracle("port").load(_oracle_tables(state["extra"])).dns_answers(blob, offs):
            state["extra"][int(r["ip"])] = (int(r["domain_hash"]), 0x7FFFFFFF)
    _check_scenarios(classify, apply_updates)
    # TestFirewall_Bypass (firewall_test.go:101-135): the same ClientHello before / during / after the bypass
    p = FX["bypass"]["packets"][0]
    blob, offs = _frames([p])
    for cg_bypassed, ex in ((False, p["expect_enforced"]), (True, p["expect_bypassed"]), (False, p["expect_enforced"])):
        t = _oracle_tables()
        if cg_bypassed:
            t["bypass"] = {**t["bypass"], 1001: 1}
        f = abi.verdict_fields(oracle.Oracle("port").load(t).classify(blob, offs, flags=F)[0])
        assert {k: int(f[k][0]) for k in ex} == ex


@pytest.mark.gpu
def test_gpu_agrees_with_third_party_decoders():
    from clawker_b200 import engine
    blob, offs = _frames(FX["cases"])
    e = engine.Engine(max_metrics=1 << 16)
    e.load_tables(capture.base_tables(8))
    e.compile_rules(RULES)
    e.publish()
    got = e.classify_varlen(blob, offs, flags=F)                      # chunk kernel: parsers read the frames from shared memory
    _check_cases(got, e.dns_answers(blob, offs, got))
    stride = 1536                                                     # the same frames as fixed slots: generic kernel, parsers read global memory
    slots = np.zeros((len(FX["cases"]), stride), np.uint8)
    for i, c in enumerate(FX["cases"]):
        fr = bytes.fromhex(c["frame"])
        slots[i, :len(fr)] = np.frombuffer(fr, np.uint8)
    gs = e.classify_slots(slots, stride=stride, flags=F)
    f, g = abi.verdict_fields(got), abi.verdict_fields(gs)
    assert np.array_equal(f["name"], g["name"]) and np.array_equal(f["nupd"], g["nupd"])
    # large batch: every fixture frame many times over, shuffled, so the chunk kernel's lists and units are full
    rng = np.random.Generator(np.random.PCG64(5))
    idx = rng.integers(0, len(FX["cases"]), size=60_000)
    frames = [bytes.fromhex(c["frame"]) for c in FX["cases"]]
    b2, o2 = capture.build_varlen([frames[i] for i in idx])
    g2 = e.classify_varlen(b2, o2, flags=F)
    assert np.array_equal(g2, got[idx])


@pytest.mark.gpu
def test_gpu_replays_the_reference_e2e_expectations():
    from clawker_b200 import engine
    e = engine.Engine(max_metrics=1 << 16)
    e.load_tables(capture.base_tables(8))
    e.compile_rules(RULES)
    e.publish()
    _check_scenarios(lambda blob, offs: e.classify_varlen(blob, offs, flags=F),
                     lambda blob, offs, w: e.dns_answers(blob, offs, w, apply=True, now=1000))
    p = FX["bypass"]["packets"][0]
    blob, offs = _frames([p])
    for bypassed, ex in ((False, p["expect_enforced"]), (True, p["expect_bypassed"]), (False, p["expect_enforced"])):
        if bypassed:
            e.map_update(abi.MAP_BYPASS, C.c_uint64(1001), C.c_uint8(1))
        else:
            e.map_delete(abi.MAP_BYPASS, C.c_uint64(1001))
        f = abi.verdict_fields(e.classify_varlen(blob, offs, flags=F))
        assert {k: int(f[k][0]) for k in ex} == ex
