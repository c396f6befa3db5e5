"""Compact ingest (clk_hdr16): the 16-byte record carries exactly what the decision reads of an IPv4 header-only frame, so
verdicts and metrics through clk_classify_hdr16 equal clk_classify_slots on the same frames (and the oracle)."""
import numpy as np
import pytest

from clawker_b200 import abi, capture, engine


def _frames(n=50_000):
    w = capture.make_workload("c2")
    return w, capture.gen_slots(w, n, seed=9, v6_frac=0.03, nonip_frac=0.01)


def test_packer_fields_and_kinds():
    w, pk = _frames(20_000)
    recs, aux = engine.pack_hdr16(pk)
    assert recs.shape[0] == pk.shape[0] and aux.size % 64 == 0
    et = (pk[:, 12].astype(np.uint32) << 8) | pk[:, 13]
    v4 = (et == 0x0800) & (pk[:, 14] == 0x45) & (((pk[:, 20] & 0x1F) | pk[:, 21]) == 0)
    nonip = (et != 0x0800) & (et != 0x86DD)
    assert np.array_equal(recs["kind"] == 0, v4) and np.array_equal(recs["kind"] == 1, nonip)
    k2 = np.flatnonzero(recs["kind"] == 2)
    assert k2.size == aux.size // 64 and k2.size > 0
    assert np.array_equal(aux.reshape(-1, 64)[recs["src"][k2]], pk[k2, :64])            # whole frames travel unchanged
    i = np.flatnonzero(v4)
    assert np.array_equal(recs["src"][i], pk[i, 26:30].copy().view("<u4").ravel()) and np.array_equal(recs["dst"][i], pk[i, 30:34].copy().view("<u4").ravel())
    assert np.array_equal(recs["dport"][i], (pk[i, 36].astype(np.uint16) << 8) | pk[i, 37]) and np.array_equal(recs["proto"][i], pk[i, 23])


def test_packer_reports_a_full_aux():
    import ctypes as C
    w, pk = _frames(4_000)
    recs = np.zeros(pk.shape[0], dtype=abi.HDR16_DTYPE)
    aux, n_aux = np.zeros(64, np.uint8), C.c_uint64()
    assert engine.lib().clk_pack_hdr16(pk.ctypes.data, 64, pk.shape[0], recs.ctypes.data, aux.ctypes.data, 1, C.byref(n_aux)) == abi.E2BIG


@pytest.mark.gpu
def test_hdr16_equals_slots_and_oracle():
    from util import assert_same_metrics, assert_same_verdicts, make_engine, oracle_tables
    from oracle import oracle
    w, pk = _frames(300_000)
    want, wm = oracle.Oracle("port").load(oracle_tables(w)).classify(pk)
    e = make_engine(w)
    got_slots = e.classify_slots(pk)
    assert_same_verdicts(got_slots, want, "slots")
    e.counters_reset()
    recs, aux = engine.pack_hdr16(pk)
    got = e.classify_hdr16(recs, aux)
    assert_same_verdicts(got, want, "hdr16")
    assert_same_metrics(e.counters(), wm, "hdr16 metrics")
    # pinned records through the same call (the e2e form), and an aux-free batch
    recs_p, aux_p = engine.pack_hdr16(pk, pinned=True)
    assert_same_verdicts(e.classify_hdr16(recs_p, aux_p), want, "hdr16 pinned")
    v4 = np.flatnonzero(recs["kind"] != 2)
    assert_same_verdicts(e.classify_hdr16(np.ascontiguousarray(recs[v4]), None), want[v4], "hdr16 without aux")
