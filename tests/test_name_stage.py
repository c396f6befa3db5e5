"""Name-stage edge cases: expectations pinned on the oracle (CPU), parity on the GPU."""
import numpy as np
import pytest

from clawker_b200 import abi, capture
from oracle import oracle, rules_oracle

import name_cases as NC


def _tables():
    t = capture.base_tables(8)
    c = rules_oracle.compile_rules(NC.RULES)
    t.update(zones=c.zones, snis=c.snis, routes=dict(c.routes), cidrs=list(c.cidrs))
    return t


def _batch():
    return capture.build_varlen([fr for _, fr, _, _ in NC.CASES])


def test_oracle_matches_documented_expectations():
    blob, offs = _batch()
    v, _ = oracle.Oracle("port").load(_tables()).classify(blob, offs, flags=abi.F_NAME_STAGE)
    f = abi.verdict_fields(v)
    for i, (desc, _, name, nupd) in enumerate(NC.CASES):
        assert int(f["name"][i]) == name, f"{desc}: name verdict {int(f['name'][i])}, expected {name}"
        if nupd is not None:
            assert int(f["nupd"][i]) == nupd, f"{desc}: {int(f['nupd'][i])} updates, expected {nupd}"


def test_oracle_update_stream_contents():
    blob, offs = _batch()
    upd = oracle.Oracle("port").load(_tables()).dns_answers(blob, offs)
    idx = {d: i for i, (d, _, _, _) in enumerate(NC.CASES)}
    one = upd[upd["packet"] == idx["one A record"]]
    assert one.size == 1 and int(one["ip"][0]) == capture.ip4("140.82.121.4") and int(one["ttl"][0]) == 300
    assert int(one["domain_hash"][0]) == rules_oracle.domain_hash("github.com")
    two = upd[upd["packet"] == idx["two A + AAAA + CNAME"]]
    assert two.size == 2 and list(two["ttl"]) == [60, 60]                      # TTL 0 -> 60 (dnsbpf.go:91-94)
    assert int(two["domain_hash"][0]) == rules_oracle.domain_hash("example.org")   # hash of the ZONE, not of the query name
    assert (upd["packet"] == idx["40 A records: capped at 31"]).sum() == 31
    assert np.all(np.diff(upd["packet"].astype(np.int64)) >= 0)                # packet order


@pytest.mark.gpu
def test_gpu_matches_oracle_on_name_cases():
    from util import assert_same_verdicts
    from clawker_b200 import engine
    blob, offs = _batch()
    o = oracle.Oracle("port").load(_tables())
    want, _ = o.classify(blob, offs, flags=abi.F_NAME_STAGE)
    e = engine.Engine(max_metrics=1 << 16)
    e.load_tables(capture.base_tables(8))
    e.compile_rules(NC.RULES)
    e.publish()
    got = e.classify_varlen(blob, offs, flags=abi.F_NAME_STAGE)
    assert_same_verdicts(got, want, "name cases")
    a, b = e.dns_answers(blob, offs, got), o.dns_answers(blob, offs)
    assert a.size == b.size and np.array_equal(a, b)
    # the same frames as fixed 768-byte slots take the generic kernel: same words
    stride = 768
    slots = np.zeros((len(NC.CASES), stride), np.uint8)
    for i, (_, fr, _, _) in enumerate(NC.CASES):
        slots[i, :len(fr)] = np.frombuffer(fr, np.uint8)
    want_s, _ = o.classify(slots, None, stride=stride, flags=abi.F_NAME_STAGE)
    assert_same_verdicts(e.classify_slots(slots, stride=stride, flags=abi.F_NAME_STAGE), want_s, "name cases in slots")
    # applying the update stream feeds dns_cache: a later TCP connect to a resolved address is attributed to the zone
    e.dns_answers(blob, offs, got, apply=True, now=1000)
    v = abi.DNSEntry()
    import ctypes as C
    assert e.map_lookup(abi.MAP_DNS_CACHE, C.c_uint32(capture.ip4("140.82.121.4")), v) is not None
    # last write in packet order wins (UpdateAny): the later "0.0.0.0 is skipped" answer re-resolves it with TTL 60
    assert v.domain_hash == rules_oracle.domain_hash("github.com") and v.expire_ts == 1060


def _mutated_batch(seed: int, n: int = 6000):
    """Name-stage frames with random byte / length corruption (parsers must agree on every malformed input)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    base = [fr for _, fr, _, _ in NC.CASES]
    frames = []
    for _ in range(n):
        fr = bytearray(base[int(rng.integers(0, len(base)))])
        k = int(rng.integers(0, 4))
        for _m in range(k):
            op = int(rng.integers(0, 4))
            lo = 34 if len(fr) > 60 else 12                   # mostly hit L4 header + payload, sometimes L3
            pos = int(rng.integers(lo, len(fr)))
            if op == 0:
                fr[pos] = int(rng.integers(0, 256))
            elif op == 1:
                fr[pos] ^= 1 << int(rng.integers(0, 8))
            elif op == 2 and len(fr) > 60:
                del fr[int(rng.integers(54, len(fr))):]      # truncate the payload (IP total length now lies)
            else:
                fr[pos:pos] = bytes(int(x) for x in rng.integers(0, 256, size=int(rng.integers(1, 9))))
        frames.append(bytes(fr))
    return capture.build_varlen(frames)


def test_oracle_survives_mutations():
    blob, offs = _mutated_batch(1, 3000)
    v, _ = oracle.Oracle("port").load(_tables()).classify(blob, offs, flags=abi.F_NAME_STAGE)
    assert v.size == offs.size - 1


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [11, 12, 13])
def test_gpu_matches_oracle_on_mutated_frames(seed):
    from util import assert_same_verdicts
    from clawker_b200 import engine
    blob, offs = _mutated_batch(seed)
    o = oracle.Oracle("port").load(_tables())
    want, wm = o.classify(blob, offs, flags=abi.F_NAME_STAGE)
    e = engine.Engine(max_metrics=1 << 16)
    e.load_tables(capture.base_tables(8))
    e.compile_rules(NC.RULES)
    e.publish()
    got = e.classify_varlen(blob, offs, flags=abi.F_NAME_STAGE)
    assert_same_verdicts(got, want, f"mutated frames seed {seed}")
    a, b = e.dns_answers(blob, offs, got), o.dns_answers(blob, offs)
    assert a.size == b.size and np.array_equal(a, b)
