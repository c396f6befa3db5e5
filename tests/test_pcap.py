"""Capture-file ingest: classic pcap -> varlen layout, natively (clk_pcap_pack), equal to the
Python statement capture.build_varlen on the same frames; byte order, nanosecond magic, truncation, chunking."""
import struct

import numpy as np
import pytest

from clawker_b200 import capture, engine
from name_cases import CASES


def _frames():
    fr = [c[1] if isinstance(c, tuple) else c for c in CASES]
    fr = [f for f in fr if isinstance(f, (bytes, bytearray))]
    assert len(fr) > 20
    return [bytes(f) for f in fr] + [b"\x01" * 15, b"\x02" * 64, b"\x03" * 65, b"\x04" * 1514]


def _write(path, frames, fmt="<", magic=0xA1B2C3D4, link=1):
    with open(path, "wb") as fh:
        fh.write(struct.pack(fmt + "IHHiIII", magic, 2, 4, 0, 0, 65535, link))
        for i, f in enumerate(frames):
            fh.write(struct.pack(fmt + "IIII", i, 0, len(f), len(f)) + f)


def _all(path, **kw):
    chunks = list(engine.pcap_pack(str(path), **kw))
    blob = np.concatenate([b for b, _ in chunks]) if chunks else np.zeros(0, dtype=np.uint8)
    offs, base = [np.zeros(1, dtype=np.int64)], 0
    for b, o in chunks:
        offs.append(o[1:].astype(np.int64) + base)
        base += b.size // 16
    return blob, np.concatenate(offs).astype(np.uint32), len(chunks)


@pytest.mark.parametrize("fmt,magic", [("<", 0xA1B2C3D4), (">", 0xA1B2C3D4), ("<", 0xA1B23C4D), (">", 0xA1B23C4D)])
def test_pack_equals_build_varlen(tmp_path, fmt, magic):
    engine.build_library()
    fr = _frames()
    _write(tmp_path / "a.pcap", fr, fmt, magic)
    wb, wo = capture.build_varlen(fr)
    blob, offs, _ = _all(tmp_path / "a.pcap")
    assert np.array_equal(offs, wo) and np.array_equal(blob, wb)
    blob, offs, n_chunks = _all(tmp_path / "a.pcap", cap_bytes=4096, cap_frames=7)     # many chunks, same result
    assert n_chunks > 5 and np.array_equal(offs, wo) and np.array_equal(blob, wb)


def test_python_reader_agrees(tmp_path):
    fr = _frames()
    capture.write_pcap(str(tmp_path / "b.pcap"), fr)
    rb, ro = capture.read_pcap(str(tmp_path / "b.pcap"))
    blob, offs, _ = _all(tmp_path / "b.pcap")
    assert np.array_equal(rb, blob) and np.array_equal(ro, offs)


def test_truncated_and_bad_files(tmp_path):
    fr = _frames()
    _write(tmp_path / "t.pcap", fr)
    raw = open(tmp_path / "t.pcap", "rb").read()
    open(tmp_path / "t.pcap", "wb").write(raw[:-100])                   # last record cut short: the file ends before it
    wb, wo = capture.build_varlen(fr[:-1])
    blob, offs, _ = _all(tmp_path / "t.pcap")
    assert np.array_equal(offs, wo) and np.array_equal(blob, wb)
    open(tmp_path / "e.pcap", "wb").write(raw[:24])                     # header only
    assert _all(tmp_path / "e.pcap")[1].tolist() == [0]
    for name, data, msg in (("m.pcap", b"\x00" * 64, "magic"), ("s.pcap", raw[:10], "shorter"),
                            ("l.pcap", struct.pack("<IHHiIII", 0xA1B2C3D4, 2, 4, 0, 0, 65535, 113) + raw[24:], "not Ethernet")):
        open(tmp_path / name, "wb").write(data)
        with pytest.raises(engine.ClkError, match=msg):
            _all(tmp_path / name)
    with pytest.raises(engine.ClkError):
        _all(tmp_path / "missing.pcap")
    _write(tmp_path / "big.pcap", fr)
    with pytest.raises(engine.ClkError, match="does not fit"):
        _all(tmp_path / "big.pcap", cap_bytes=1024)                     # the 1514-byte frame needs 1520


@pytest.mark.gpu
def test_classify_pcap_equals_classify_varlen(tmp_path):
    from oracle import oracle
    from util import assert_same_verdicts, make_engine, oracle_tables
    w = capture.make_workload("c3", n_rules=300)
    blob, offs, _ = capture.gen_varlen(w, 60_000, seed=5, pool=512)
    frames = [bytes(blob[int(offs[i]) * 16:int(offs[i + 1]) * 16]) for i in range(offs.size - 1)]
    capture.write_pcap(str(tmp_path / "c3.pcap"), frames)
    e = make_engine(w)
    want, _ = oracle.Oracle("port").load(oracle_tables(w)).classify(blob, offs, flags=1)
    assert_same_verdicts(e.classify_pcap(str(tmp_path / "c3.pcap"), flags=1), want, "pcap file")
    assert_same_verdicts(e.classify_varlen(blob, offs, flags=1), want, "same frames, host buffers")


def test_pack_survives_corruption(tmp_path):
    """Random truncations and byte flips of a valid capture: frames or an error, never a crash; offsets stay monotonic."""
    engine.build_library()
    fr = _frames()
    _write(tmp_path / "ok.pcap", fr)
    raw = open(tmp_path / "ok.pcap", "rb").read()
    rng = np.random.Generator(np.random.PCG64(9))
    for _ in range(200):
        b = bytearray(raw)
        if rng.random() < 0.5:
            b = b[: int(rng.integers(0, len(b)))]
        for _k in range(int(rng.integers(0, 8))):
            if b:
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        open(tmp_path / "fz.pcap", "wb").write(bytes(b))
        try:
            blob, offs, _ = _all(tmp_path / "fz.pcap", cap_bytes=1 << 20, cap_frames=4096)
        except engine.ClkError:
            continue
        assert np.all(np.diff(offs.astype(np.int64)) >= 4) and int(offs[-1]) * 16 == blob.size
