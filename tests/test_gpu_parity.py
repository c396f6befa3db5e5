"""GPU parity: the CUDA path (through the C-ABI) vs the CPU oracle, bit for bit."""
import numpy as np
import pytest

from clawker_b200 import abi, capture
from oracle import oracle

from util import assert_same_metrics, assert_same_verdicts, make_engine, oracle_tables

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("loader", [0, 1, 2])
def test_c2_slots_all_loaders(loader):
    w = capture.make_workload("c2")
    pk = capture.gen_slots(w, 300_001, seed=2)
    want, wm = oracle.Oracle("port").load(oracle_tables(w)).classify(pk)
    e = make_engine(w, loader=loader)
    got = e.classify_slots(pk)
    assert_same_verdicts(got, want, f"loader {loader}")
    assert_same_metrics(e.counters(), wm, f"loader {loader}")
    assert e.stats().metrics_dropped == 0


def test_c2_vs_reference_object():
    """No CIDR rules -> the decision stage must equal the reference's own eBPF C."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    w = capture.make_workload("c2")
    pk = capture.gen_slots(w, 200_000, seed=7)
    want, wm = oracle.Oracle("ref").load(oracle_tables(w, with_cidrs=False)).classify(pk)
    e = make_engine(w, with_cidrs=False)
    assert_same_verdicts(e.classify_slots(pk), want, "vs reference object")
    assert_same_metrics(e.counters(), wm, "vs reference object")


def test_c1_tiny():
    w = capture.make_workload("c1")
    pk = capture.gen_slots(w, 1000, seed=1)
    want, wm = oracle.Oracle("port").load(oracle_tables(w)).classify(pk)
    e = make_engine(w)
    assert_same_verdicts(e.classify_slots(pk), want, "c1")
    assert_same_metrics(e.counters(), wm, "c1")


def test_events():
    w = capture.make_workload("c2")
    ev = capture.gen_events(w, 100_000, seed=3)
    want, wm = oracle.Oracle("port").load(oracle_tables(w)).classify_events(ev)
    e = make_engine(w)
    assert_same_verdicts(e.classify_events(ev), want, "events")
    assert_same_metrics(e.counters(), wm, "events")


def test_empty_and_ragged():
    w = capture.make_workload("c1")
    e = make_engine(w)
    assert e.classify_slots(np.zeros((0, 64), np.uint8)).size == 0
    for n in (1, 31, 32, 33, 255, 256, 257, 513):
        pk = capture.gen_slots(w, n, seed=n)
        want, _ = oracle.Oracle("port").load(oracle_tables(w)).classify(pk)
        assert_same_verdicts(e.classify_slots(pk), want, f"n={n}")


def test_stride_128_slots():
    w = capture.make_workload("c2")
    pk64 = capture.gen_slots(w, 10_000, seed=5)
    pk = np.zeros((10_000, 128), np.uint8)
    pk[:, :64] = pk64
    want, _ = oracle.Oracle("port").load(oracle_tables(w)).classify(pk, stride=128)
    e = make_engine(w)
    assert_same_verdicts(e.classify_slots(pk, stride=128), want, "stride 128")


def test_varlen_name_stage():
    w = capture.make_workload("c3", n_rules=2000)
    blob, offs, info = capture.gen_varlen(w, 60_000, seed=3, pool=1024)
    o = oracle.Oracle("port").load(oracle_tables(w))
    want, wm = o.classify(blob, offs, flags=abi.F_NAME_STAGE)
    e = make_engine(w)
    got = e.classify_varlen(blob, offs, flags=abi.F_NAME_STAGE)
    assert_same_verdicts(got, want, "c3 varlen")
    assert_same_metrics(e.counters(), wm, "c3 varlen")
    upd_want = o.dns_answers(blob, offs)
    upd_got = e.dns_answers(blob, offs, got)
    assert upd_got.size == upd_want.size and np.array_equal(upd_got, upd_want)
    # sanity: the capture really exercises every stage
    f = abi.verdict_fields(got)
    assert (f["name"] == 1).sum() > 100 and (f["name"] == 2).sum() > 100 and f["nupd"].sum() > 100
