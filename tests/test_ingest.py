"""The capture front end either side of the classifier: TCP reassembly of split ClientHellos / HTTP request heads
(clk_tcp_coalesce) and live AF_PACKET ingest (clk_afpacket_*).  The expectation for a reassembled record is the oracle's verdict
for the same bytes in ONE frame; the big ClientHello comes from OpenSSL (ssl.MemoryBIO), as in tests/golden/make_wire_fixtures.py."""
import json
import socket
import ssl
import time
from pathlib import Path

import numpy as np
import pytest

from clawker_b200 import abi, capture, engine
from clawker_b200.capture import frame, ip4
from oracle import oracle, rules_oracle

FX = json.loads((Path(__file__).resolve().parent / "golden" / "wire_fixtures.json").read_text())
RULES = [{**r, "path_rules": [tuple(p) for p in r.get("path_rules", [])]} for r in FX["rules"]]
F = abi.F_NAME_STAGE
CT, EXT = ip4("172.20.0.10"), ip4("93.184.216.34")
NA, ALLOW, DENY = 0, 1, 2


def _tables():
    t = capture.base_tables(8)
    c = rules_oracle.compile_rules(RULES)
    t.update(zones=c.zones, snis=c.snis, routes=dict(c.routes), cidrs=list(c.cidrs), http=list(c.http))
    return t


def _orc(blob, offs, flags=F):
    return oracle.Oracle("port").load(_tables()).classify(blob, offs, flags=flags)[0]


def _hello(host, n_alpn=40):
    ctx = ssl.SSLContext(ssl.PROTOCOL_TLS_CLIENT)
    ctx.check_hostname, ctx.verify_mode = False, ssl.CERT_NONE
    ctx.set_alpn_protocols(["proto-%03d-%s" % (i, "x" * 40) for i in range(n_alpn)])
    inc, out = ssl.MemoryBIO(), ssl.MemoryBIO()
    try:
        ctx.wrap_bio(inc, out, server_hostname=host).do_handshake()
    except ssl.SSLWantReadError:
        pass
    return out.read()


def _segments(src, dst, sport, dport, payload, seq0, mss=1460):
    return [frame(src, dst, 6, sport, dport, payload[o:o + mss], seq=seq0 + o) for o in range(0, len(payload), mss)]


def _payload(blob, offs, i):
    p = bytes(blob[int(offs[i]) * 16: int(offs[i + 1]) * 16])
    end = 14 + int.from_bytes(p[16:18], "big")
    return p[54:end]


def _scenario():
    """Interleaved flows; returns (frames, expected payload per frame after reassembly, expected name verdict per frame or None)."""
    big_ok, big_no = _hello("api.anthropic.com"), _hello("evil.example.net", 70)
    assert len(big_ok) > 1460 and len(big_no) > 2920
    http = (b"GET /test/page HTTP/1.1\r\nX-Filler: " + b"f" * 1500 + b"\r\nHost: example.com\r\nAccept: */*\r\n\r\n")
    small = _hello("sentry.io", 1)
    a, b, c = _segments(CT, EXT, 40001, 443, big_ok, 1000), _segments(CT, EXT, 40002, 443, big_no, 5000), _segments(CT, EXT, 40003, 80, http, 77)
    assert len(a) == 2 and len(b) == 3 and len(c) == 2
    ooo = _segments(CT, EXT, 40004, 443, big_ok, 9000)                       # second segment arrives with a gap: never merged
    ooo[1] = frame(CT, EXT, 6, 40004, 443, big_ok[1460:], seq=9000 + 1460 + 7)
    rst = _segments(CT, EXT, 40005, 443, big_ok, 100)
    wrap = _segments(CT, EXT, 40006, 443, big_ok, 0xFFFFFFFF - 500)          # sequence numbers wrap inside the hello
    dnsq = frame(CT, ip4("127.0.0.11"), 17, 40000, 53, capture.dns_query("github.com"))
    frames = [frame(CT, EXT, 6, 40001, 443, b"", tcp_flags=0x02, seq=999), a[0], dnsq, b[0], a[1], c[0], b[1], frame(CT, EXT, 6, 40007, 443, small),
              c[1], b[2], ooo[0], ooo[1], rst[0], frame(CT, EXT, 6, 40005, 443, b"", tcp_flags=0x04, seq=100 + 1460), rst[1], wrap[0], wrap[1],
              frame(CT, EXT, 6, 40001, 443, b"\x17\x03\x03\x00\x10" + b"d" * 16, seq=1000 + len(big_ok))]
    whole = {1: big_ok, 3: big_no, 5: http, 15: big_ok}
    empty = {4, 6, 8, 9, 16}
    return frames, whole, empty


def test_coalesce_layout_and_payloads():
    frames, whole, empty = _scenario()
    blob, offs = capture.build_varlen(frames)
    out, oo, merged = engine.tcp_coalesce(blob, offs)
    assert merged == len(empty) and oo.size == offs.size and int(oo[-1]) * 16 == out.size
    for i, fr in enumerate(frames):
        got = _payload(out, oo, i)
        if i in whole:
            assert got == whole[i], i
        elif i in empty:
            assert got == b"", i
        else:
            assert got == _payload(blob, offs, i) and bytes(out[int(oo[i]) * 16:int(oo[i]) * 16 + len(fr)]) == fr, i
        assert (int(oo[i + 1]) - int(oo[i])) * 16 >= 64


def _expected_names(frames, whole, empty):
    """The oracle's verdict for each record in ONE frame (what Envoy decides once it has buffered the record)."""
    single = []
    for i, fr in enumerate(frames):
        single.append(fr[:54] if i in empty else frame(CT, EXT, 6, int.from_bytes(fr[34:36], "big"), int.from_bytes(fr[36:38], "big"), whole[i],
                                                        seq=int.from_bytes(fr[38:42], "big")) if i in whole else fr)
    b, o = capture.build_varlen(single)
    return _orc(b, o)


def test_coalesced_batch_classifies_like_whole_records():
    frames, whole, empty = _scenario()
    blob, offs = capture.build_varlen(frames)
    before = _orc(blob, offs)
    fb = abi.verdict_fields(before)
    assert all(int(fb["name"][i]) == NA for i in whole)                      # first segments alone: needs reassembly
    out, oo, _ = engine.tcp_coalesce(blob, offs)
    after = _orc(out, oo)
    want = _expected_names(frames, whole, empty)
    assert np.array_equal(after, want)
    fa = abi.verdict_fields(after)
    assert [int(fa["name"][i]) for i in (1, 3, 5, 15)] == [ALLOW, DENY, ALLOW, ALLOW]
    assert all(int(fa["name"][i]) == NA for i in empty)
    assert int(fa["name"][10]) == NA and int(fa["name"][12]) == NA           # gap / reset flows stay undecided
    assert int(fa["name"][7]) == int(fb["name"][7]) == ALLOW                 # complete records are left alone


def test_coalesce_edges():
    blob, offs = capture.build_varlen([])
    out, oo, merged = engine.tcp_coalesce(blob, offs)
    assert merged == 0 and oo.tolist() == [0]
    frames = [frame(CT, EXT, 17, 1, 2, b"x" * 10), b"\x00" * 20, frame(CT, EXT, 6, 5, 443, _hello("api.anthropic.com")[:1460], seq=1)]
    blob, offs = capture.build_varlen(frames)
    out, oo, merged = engine.tcp_coalesce(blob, offs)                        # nothing to merge: byte-identical copy
    assert merged == 0 and np.array_equal(oo, offs) and np.array_equal(out, blob)
    L = engine.lib()
    n = np.zeros(1, np.uint64)
    small = np.zeros(64, np.uint8)
    rc = L.clk_tcp_coalesce(blob.ctypes.data, offs.ctypes.data, 3, small.ctypes.data, small.size, np.zeros(4, np.uint32).ctypes.data,
                            n.ctypes.data_as(engine.C.POINTER(engine.C.c_uint64)))
    assert rc == abi.E2BIG and b"too small" in L.clk_ingest_last_error()


def test_afpacket_ring_on_loopback():
    try:
        ring = engine.PacketRing("lo", ring_mib=4)
    except engine.ClkError as x:
        pytest.skip(f"AF_PACKET not available here: {x}")
    with pytest.raises(engine.ClkError) as ei:
        engine.PacketRing("no-such-if0")
    assert ei.value.code == abi.ENOENT
    with ring:
        s = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        tag = b"clawker-ring-%d-" % time.time_ns()
        sent = [tag + b"%04d" % i + b"p" * (i * 7 % 200) for i in range(300)]
        for m in sent:
            s.sendto(m, ("127.0.0.1", 40053))
        s.close()
        seen, deadline = [], time.time() + 5
        while len(seen) < len(sent) and time.time() < deadline:
            blob, offs = ring.read(cap_bytes=1 << 16, cap_frames=64, timeout_ms=200)      # small buffers: blocks are drained across calls
            assert offs.size <= 65 and int(offs[-1]) * 16 == blob.size
            for i in range(offs.size - 1):
                fr = bytes(blob[int(offs[i]) * 16:int(offs[i + 1]) * 16])
                assert len(fr) >= 64 and len(fr) % 16 == 0
                if fr[12:14] == b"\x08\x00" and fr[23] == 17 and fr[42:42 + len(tag)] == tag:
                    seen.append(fr[42:14 + int.from_bytes(fr[16:18], "big")])
        assert seen == sent                                                  # every datagram once, in order, bytes intact
        s = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        s.sendto(b"x" * 1000, ("127.0.0.1", 40053))
        s.close()
        time.sleep(0.05)
        with pytest.raises(engine.ClkError) as big:                          # a frame that fits no buffer is an error, not an endless n == 0
            ring.read(cap_bytes=256, cap_frames=8, timeout_ms=200)
        assert big.value.code == abi.E2BIG
        assert ring.read(cap_bytes=1 << 16, cap_frames=8, timeout_ms=200)[1].size >= 2       # ... and the frame is still there for a larger one
        w = _orc(blob, offs, 0)                                              # and the chunks are valid classifier input
        assert w.size == offs.size - 1


def _interleaved(n_flows=200):
    """The scenario n_flows times over on distinct ports, the flows interleaved at random (3600 frames: inside the 4096-frame
    window after which a flow that went quiet is given up)."""
    frames, whole, empty = _scenario()
    rng = np.random.Generator(np.random.PCG64(3))
    flows = [[fr[:34] + (int.from_bytes(fr[34:36], "big") + 10 * k).to_bytes(2, "big") + fr[36:] if fr[23] == 6 else fr for fr in frames]
             for k in range(n_flows)]
    order = np.repeat(np.arange(n_flows), len(frames))
    rng.shuffle(order)
    nxt, batch = [0] * n_flows, []
    for k in order:
        batch.append(flows[k][nxt[k]])
        nxt[k] += 1
    return batch, n_flows * len(empty)


def test_coalesce_interleaved_flows():
    batch, want_merged = _interleaved()
    blob, offs = capture.build_varlen(batch)
    out, oo, merged = engine.tcp_coalesce(blob, offs)
    assert merged == want_merged
    names = abi.verdict_fields(_orc(out, oo))["name"]
    assert int((names == ALLOW).sum()) == 200 * 5 and int((names == DENY).sum()) == 200          # per scenario: 3 hellos + HTTP + DNS query allowed, 1 hello denied


@pytest.mark.gpu
def test_gpu_classifies_the_coalesced_batch():
    batch, want_merged = _interleaved()
    blob, offs = capture.build_varlen(batch)
    out, oo, merged = engine.tcp_coalesce(blob, offs)
    assert merged == want_merged
    e = engine.Engine(max_metrics=1 << 16)
    e.load_tables(capture.base_tables(8))
    e.compile_rules(RULES)
    e.publish()
    got = e.classify_varlen(out, oo, flags=F)
    assert np.array_equal(got, _orc(out, oo))
    e.close()


def test_afpacket_ring_to_hdr16_records():
    """The ring drained straight into compact ingest records: same records as packing the captured frames (clk_pack_hdr16)."""
    try:
        ring = engine.PacketRing("lo", ring_mib=4)
        ring2 = engine.PacketRing("lo", ring_mib=4)
    except engine.ClkError as x:
        pytest.skip(f"AF_PACKET not available here: {x}")
    with ring, ring2:
        s = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        s6 = socket.socket(socket.AF_INET6, socket.SOCK_DGRAM)
        tag = b"clawker-h16-%d-" % time.time_ns()
        for i in range(200):
            s.sendto(tag + b"%04d" % i, ("127.0.0.1", 40054))
            if i % 10 == 0:
                try:
                    s6.sendto(tag + b"v6%02d" % i, ("::1", 40054))      # IPv6 frames: aux slots
                except OSError:
                    pass
        s.close(); s6.close()
        time.sleep(0.1)
        frames, recs, auxs, deadline = [], [], [], time.time() + 5
        while time.time() < deadline:
            blob, offs = ring.read(cap_bytes=1 << 20, cap_frames=4096, timeout_ms=100)
            r, a = ring2.read_hdr16(cap_frames=4096, aux_cap=64, timeout_ms=100)
            frames += [bytes(blob[int(offs[i]) * 16:int(offs[i + 1]) * 16]) for i in range(offs.size - 1)]
            base = sum(x.size // 64 for x in auxs)
            r = r.copy()
            r["src"][r["kind"] == 2] += base                                   # CLK_HDR_AUX: slot index within this call -> within the run
            recs.append(r); auxs.append(a)
            if offs.size == 1 and r.size == 0 and frames:
                break
        recs, aux = np.concatenate(recs), np.concatenate(auxs) if auxs else np.zeros(0, np.uint8)
        # the two rings may differ by unrelated loopback traffic around their opening: compare what this test sent
        mine = [i for i, fr in enumerate(frames) if len(fr) >= 64 and fr[12:14] == b"\x08\x00" and fr[23] == 17 and fr[36:38] == (40054).to_bytes(2, "big")]
        assert len(mine) == 200
        slots = np.zeros((len(mine), 64), np.uint8)
        for k, i in enumerate(mine):
            slots[k] = np.frombuffer(frames[i][:64], np.uint8)
        want, _ = engine.pack_hdr16(slots.reshape(-1), 64)                    # the host packer on the same frames
        got = recs[(recs["kind"] == 0) & (recs["proto"] == 17) & (recs["dport"] == 40054)]
        assert np.array_equal(got, want)
        n6 = sum(1 for fr in frames if fr[12:14] == b"\x86\xdd")
        assert int((recs["kind"] == 2).sum()) == aux.size // 64 >= min(n6, 1)
        for j in np.flatnonzero(recs["kind"] == 2)[:8]:                       # an aux slot holds the frame's first 64 bytes
            assert bytes(aux[int(recs["src"][j]) * 64:][:14]) [12:14] in (b"\x86\xdd", b"\x08\x00")


def test_coalesce_fuzz_keeps_the_layout():
    """Random truncations, byte flips and reorderings of the scenario: never a crash, one output frame per input frame, offsets
    monotonic, frames the pass does not touch byte-identical, and the pass is idempotent on its own output's untouched frames."""
    frames, _, _ = _scenario()
    rng = np.random.Generator(np.random.PCG64(11))
    for trial in range(300):
        fs = [bytearray(f) for f in frames]
        rng.shuffle(fs)
        for f in fs:
            u = rng.random()
            if u < 0.25 and len(f) > 20:
                del f[int(rng.integers(14, len(f))):]                        # truncate anywhere after the Ethernet header
            elif u < 0.5:
                for _ in range(int(rng.integers(1, 6))):
                    f[int(rng.integers(0, len(f)))] = int(rng.integers(0, 256))
        fs = [bytes(f) for f in fs]
        blob, offs = capture.build_varlen(fs)
        out, oo, merged = engine.tcp_coalesce(blob, offs)
        assert oo.size == offs.size and np.all(np.diff(oo.astype(np.int64)) >= 4) and int(oo[-1]) * 16 == out.size and merged <= len(fs)
        untouched = 0
        for i, f in enumerate(fs):
            got = bytes(out[int(oo[i]) * 16:int(oo[i + 1]) * 16])
            if got[:len(f)] == f:
                untouched += 1
            else:
                assert got[:14] == f[:14] and got[23] == 6                   # only TCP frames are rewritten, and only past their headers
        assert len(fs) - untouched <= 2 * merged                              # a merge rewrites its first segment and the folded one, nothing else
