"""Plaintext-HTTP Host / path-prefix stage: rule compilation against the reference's own test
vectors, the oracle against documented expectations, the GPU against the oracle (cases + byte mutations)."""
import numpy as np
import pytest

from clawker_b200 import abi, capture, engine
from oracle import oracle, rules_oracle

import http_cases as HC


def _tables(rules=HC.RULES):
    t = capture.base_tables(8)
    c = rules_oracle.compile_rules(rules)
    t.update(zones=c.zones, snis=c.snis, routes=dict(c.routes), cidrs=list(c.cidrs), http=list(c.http))
    return t


def _dump(rules):
    e = engine.Engine(host_only=True)
    e.compile_rules(rules)
    out = []
    for line in e.rules_dump(3):
        d, r = line.split("|")
        out.append((d.split(","), [(x[1:], x[0] == "A") for x in r.split(",")]))
    return out


def test_http_domains_reference_vectors():
    """TestHTTPDomains_* (internal/firewall/envoy_test.go:550-572)."""
    assert rules_oracle.http_domains("api.anthropic.com", set()) == ["api.anthropic.com", "api.anthropic.com:*"]
    assert rules_oracle.http_domains(".datadoghq.com", set()) == ["*.datadoghq.com", "*.datadoghq.com:*", "datadoghq.com", "datadoghq.com:*"]
    assert rules_oracle.http_domains(".claude.ai", {"claude.ai"}) == ["*.claude.ai", "*.claude.ai:*"]
    for rules, want in (([{"dst": "api.anthropic.com", "proto": "http"}], [["api.anthropic.com", "api.anthropic.com:*"]]),
                        ([{"dst": ".datadoghq.com", "proto": "http"}], [["*.datadoghq.com", "*.datadoghq.com:*", "datadoghq.com", "datadoghq.com:*"]]),
                        ([{"dst": ".claude.ai", "proto": "http"}, {"dst": "claude.ai", "proto": "http"}], [["*.claude.ai", "*.claude.ai:*"], ["claude.ai", "claude.ai:*"]]),
                        ([{"dst": ".claude.ai", "proto": "http"}, {"dst": "claude.ai", "proto": "tls"}], [["*.claude.ai", "*.claude.ai:*", "claude.ai", "claude.ai:*"]])):
        assert [d for d, _ in _dump(rules)] == want                      # per-listener exact set (envoy.go:210-227)


def test_http_routes_reference_shape():
    """TestGenerateEnvoyConfig_HTTPWithPathRules (envoy_test.go:271-296) + buildHTTPRoutes (envoy.go:566-607)."""
    rules = [{"dst": "api.example.com", "proto": "http", "port": 80, "action": "allow",
              "path_rules": [("/api/v1", "allow"), ("/admin", "deny")], "path_default": "deny"}]
    want = [(["api.example.com", "api.example.com:*"], [("/api/v1", True), ("/admin", False), ("/", False)])]
    assert _dump(rules) == want and rules_oracle.compile_rules(rules).http == want
    for pd, allow in (("", False), ("deny", False), ("DENY", False), ("allow", True), ("Allow", True), ("whatever", True)):
        r = [dict(rules[0], path_default=pd)]
        assert _dump(r)[0][1][-1] == ("/", allow) and rules_oracle.compile_rules(r).http[0][1][-1] == ("/", allow)
    assert _dump([{"dst": "x.org", "proto": "http"}]) == [(["x.org", "x.org:*"], [("/", True)])]       # no path rules: allow all
    assert _dump([{"dst": "x.org", "proto": "http", "action": "deny"}, {"dst": "10.0.0.1", "proto": "http"}, {"dst": "x.org", "proto": "tls"}]) == []
    cxx, py = _dump(HC.RULES), rules_oracle.compile_rules(HC.RULES).http
    assert cxx == [(d, r) for d, r in py] and len(cxx) == 5


def test_oracle_matches_documented_expectations():
    blob, offs = capture.build_varlen([fr for _, fr, _ in HC.CASES])
    v, _ = oracle.Oracle("port").load(_tables()).classify(blob, offs, flags=abi.F_NAME_STAGE)
    f = abi.verdict_fields(v)
    for i, (desc, _, name) in enumerate(HC.CASES):
        assert int(f["name"][i]) == name, f"{desc}: name verdict {int(f['name'][i])}, expected {name}"
    # without HTTP rules the raw_buffer chain does not exist: every non-TLS segment is reset
    import name_cases as NC
    v2, _ = oracle.Oracle("port").load(_tables(NC.RULES)).classify(blob, offs, flags=abi.F_NAME_STAGE)
    f2 = abi.verdict_fields(v2)
    for i, (desc, fr, name) in enumerate(HC.CASES):
        if "TLS ClientHello" not in desc and desc != "empty TCP payload":
            assert int(f2["name"][i]) == HC.DENY, desc


def _mutated(seed, n=4000):
    rng = np.random.Generator(np.random.PCG64(seed))
    base = [bytearray(fr) for _, fr, _ in HC.CASES]
    out = []
    for k in range(n):
        fr = bytearray(base[int(rng.integers(0, len(base)))])
        if len(fr) > 54:
            for _ in range(int(rng.integers(1, 4))):
                i = int(rng.integers(54, len(fr)))
                fr[i] = [fr[i] ^ (1 << int(rng.integers(0, 8))), ord("\r"), ord("\n"), ord(":"), ord("."), ord(" "), ord("/"), ord("*")][int(rng.integers(0, 8))]
            if rng.random() < 0.2:
                fr = fr[: int(rng.integers(54, len(fr) + 1))]
                fr[16:18] = (len(fr) - 14).to_bytes(2, "big")            # keep the IP total length consistent with the cut
        out.append(bytes(fr))
    return out


def test_oracle_survives_mutations():
    blob, offs = capture.build_varlen(_mutated(3))
    v, _ = oracle.Oracle("port").load(_tables()).classify(blob, offs, flags=abi.F_NAME_STAGE)
    assert set(np.unique(abi.verdict_fields(v)["name"]).tolist()) <= {0, 1, 2}


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [None, 21, 22])
def test_gpu_matches_oracle(seed):
    from util import assert_same_verdicts
    frames = [fr for _, fr, _ in HC.CASES] if seed is None else _mutated(seed)
    blob, offs = capture.build_varlen(frames)
    want, _ = oracle.Oracle("port").load(_tables()).classify(blob, offs, flags=abi.F_NAME_STAGE)
    e = engine.Engine(max_metrics=1 << 16)
    e.load_tables(capture.base_tables(8))
    e.compile_rules(HC.RULES)
    e.publish()
    assert_same_verdicts(e.classify_varlen(blob, offs, flags=abi.F_NAME_STAGE), want, f"http cases (seed {seed})")


@pytest.mark.gpu
def test_gpu_matches_oracle_on_the_c3h_mix():
    """The bench workload with plaintext HTTP in the mix (bench.py --workload c3h), at a size the oracle does in seconds."""
    from util import assert_same_metrics, assert_same_verdicts, make_engine, oracle_tables
    w = capture.make_workload("c3h", n_rules=2000)
    blob, offs, info = capture.gen_varlen(w, 300_000, seed=9, pool=2048, mix=(0.64, 0.10, 0.05, 0.13, 0.08))
    assert info["classes"][4] > 20_000
    want, wm = oracle.Oracle("port").load(oracle_tables(w)).classify(blob, offs, flags=abi.F_NAME_STAGE)
    names = abi.verdict_fields(want)["name"]
    assert (names == 1).sum() > 10_000 and (names == 2).sum() > 10_000
    e = make_engine(w)
    assert_same_verdicts(e.classify_varlen(blob, offs, flags=abi.F_NAME_STAGE), want, "c3h mix")
    assert_same_metrics(e.counters(), wm, "c3h mix")
