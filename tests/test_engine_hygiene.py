"""Boundary hygiene of libclawker_gpu.so: the round-1 advisor findings, each pinned by a test."""
import ctypes as C

import numpy as np
import pytest

from clawker_b200 import abi, capture, engine


def test_state_load_keeps_dns_rows_learned_in_memory(tmp_path):
    """The control-plane verbs rewrite the state file; rows the classifying process learned since (dnsbpf's writes) must
    survive the reload, and a row both sides hold keeps the later expiry (the reference's pinned map is one shared object)."""
    p = str(tmp_path / "state.bin")
    e = engine.Engine(host_only=True)
    e.load_tables(capture.base_tables(4))
    e.dns_update(capture.ip4("140.82.121.4"), 111, ttl=60, now=1000)             # known to the file
    e.state_save(p)
    e.dns_update(capture.ip4("140.82.121.4"), 111, ttl=300, now=1000)            # refreshed in memory: later expiry
    e.dns_update(capture.ip4("1.2.3.4"), 222, ttl=60, now=1000)                  # learned in memory only
    m = engine.Engine(host_only=True)                                            # the manager: changes something else and saves
    m.state_load(p)
    m.dns_update(capture.ip4("9.9.9.9"), 333, ttl=60, now=1000)
    m.state_save(p)
    e.state_load(p)
    v = abi.DNSEntry()
    assert e.map_lookup(abi.MAP_DNS_CACHE, C.c_uint32(capture.ip4("1.2.3.4")), v) is not None and v.domain_hash == 222
    assert e.map_lookup(abi.MAP_DNS_CACHE, C.c_uint32(capture.ip4("9.9.9.9")), v) is not None and v.domain_hash == 333
    assert e.map_lookup(abi.MAP_DNS_CACHE, C.c_uint32(capture.ip4("140.82.121.4")), v) is not None and v.expire_ts == 1300


def test_comm_unique_id_or_clean_error():
    """NCCL is dlopen()ed on demand: on a box without it (or without a GPU) the call fails with a code, it never throws or aborts."""
    buf, n = C.create_string_buffer(256), C.c_size_t()
    rc = engine.lib().clk_comm_unique_id(buf, 256, C.byref(n))
    assert rc in (abi.OK, abi.ENOSYS, abi.EIO)
    assert (n.value == 128) if rc == abi.OK else True
    assert engine.lib().clk_comm_unique_id(buf, 16, C.byref(n)) == abi.EINVAL    # too small a buffer


def test_comm_calls_on_a_host_only_engine_fail_cleanly():
    e = engine.Engine(host_only=True)
    assert engine.lib().clk_allgather_counters(e.h, None) == abi.ENOSYS
    assert engine.lib().clk_comm_init(e.h, C.create_string_buffer(128), 128, 0, 1) == abi.ENOSYS


@pytest.mark.gpu
def test_more_than_eight_coredns_addresses_is_an_error():
    e = engine.Engine(max_metrics=1 << 12)
    cfg = list(capture.base_tables(1)["containers"].values())[0]
    for i in range(9):
        c = list(cfg)
        c[1] = capture.ip4(f"172.20.1.{i + 1}")                                   # coredns_ip
        e.map_update(abi.MAP_CONTAINER, C.c_uint64(5000 + i), abi.ContainerConfig(*c))
    with pytest.raises(engine.ClkError) as ei:
        e.publish()
    assert ei.value.code == abi.E2BIG and "coredns_ip" in str(ei.value)


@pytest.mark.gpu
def test_cgroup_rows_are_reclaimed_and_metrics_survive():
    """Every container start brings a new cgroup id.  Thousands of start/stop cycles must neither exhaust the cgroup-row index nor
    lose the metric rows of cgroups that still own some (round-1 advisor finding: the index only grew)."""
    from util import make_engine
    w = capture.make_workload("c2")
    e = make_engine(w, max_containers=64, max_metrics=1 << 16)
    pk = capture.gen_slots(w, 20_000, seed=3)
    e.classify_slots(pk)
    before = np.sort(e.counters(), order=["cgroup_id", "domain_hash", "dst_port", "action"])
    cfg = abi.ContainerConfig(*list(w.tables["containers"].values())[0])
    for k in range(40):                                                          # 40 x 500 short-lived containers
        for j in range(500):
            cg = C.c_uint64(100_000 + k * 500 + j)
            e.map_update(abi.MAP_CONTAINER, cg, cfg)
            if j % 50 == 49:
                e.publish()
            e.map_delete(abi.MAP_CONTAINER, cg)
        e.publish()
    after = np.sort(e.counters(), order=["cgroup_id", "domain_hash", "dst_port", "action"])
    assert np.array_equal(before, after)                                         # nothing lost, nothing double counted
    e.classify_slots(pk)                                                         # and the engine still classifies and counts
    twice = np.sort(e.counters(), order=["cgroup_id", "domain_hash", "dst_port", "action"])
    assert np.array_equal(twice["count"], 2 * before["count"])
    assert e.totals_rows() == 128                                                # container_map + bypass_map rows


@pytest.mark.gpu
def test_dev_entries_on_two_streams_do_not_share_scratch():
    """Two *_dev launches in flight on different streams of one engine (double-buffered batches): each stream has its own rare-frame
    queue / work lists, so verdicts of both batches are complete and correct."""
    import torch
    from util import make_engine, oracle_tables
    from oracle import oracle
    w = capture.make_workload("c2")
    e = make_engine(w)
    a = capture.gen_slots(w, 400_000, seed=21, v6_frac=0.2, nonip_frac=0.02)
    b = capture.gen_slots(w, 400_000, seed=22, v6_frac=0.2, nonip_frac=0.02)
    o = oracle.Oracle("port").load(oracle_tables(w))
    wa, wb = o.classify(a)[0], o.classify(b)[0]
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    va, vb = torch.zeros(a.shape[0], dtype=torch.int32, device="cuda"), torch.zeros(b.shape[0], dtype=torch.int32, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for _ in range(3):
        e.classify_slots_dev(da.data_ptr(), 64, a.shape[0], 0, va.data_ptr(), s1.cuda_stream)
        e.classify_slots_dev(db.data_ptr(), 64, b.shape[0], 0, vb.data_ptr(), s2.cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(va.cpu().numpy().view(np.uint32), wa) and np.array_equal(vb.cpu().numpy().view(np.uint32), wb)


@pytest.mark.gpu
def test_stats_before_any_timed_launch_leaves_no_pending_cuda_error():
    """clk_get_stats used to ask for the elapsed time of events that had never been recorded (host-buffer and caller-stream paths do
    not record them): an API error that stayed in the runtime's last-error slot until the next launch check (found by compute-sanitizer)."""
    import torch
    w = capture.make_workload("c2")
    pk = capture.gen_slots(w, 4096, seed=1)
    e = engine.Engine()
    e.load_tables(w.tables)
    e.publish()
    e.classify_slots(pk, 64, 0)                       # host path: no timing events
    assert e.stats().kernel_launches >= 1
    cudart = torch.cuda.cudart()
    assert int(cudart.cudaGetLastError()) == 0 if hasattr(cudart, "cudaGetLastError") else True
    e.classify_slots(pk, 64, 0)                       # and the next launch check does not trip over a stale error
    e.close()


@pytest.mark.gpu
def test_varlen_host_path_rejects_bad_offsets(monkeypatch):
    """The host-buffer varlen path copies chunk by chunk and checks a chunk's offsets while its copy is on the wire: a dip inside a
    chunk, or a chunk boundary outside the blob, is CLK_EINVAL — and nothing is read past the caller's blob."""
    monkeypatch.setenv("CLK_CHUNK_PKTS", "1024")
    w = capture.make_workload("c3")
    blob, offs = capture.gen_varlen(w, 5000, seed=4)[:2]
    e = engine.Engine()
    e.load_tables(w.tables)
    e.publish()
    good = e.classify_varlen(blob, offs, abi.F_NAME_STAGE)
    assert good.size == 5000
    for at, val in ((2500, None), (1024, 0xFFFFFFF0), (2048, 0)):
        bad = offs.copy()
        bad[at] = (bad[at + 1] + 7) if val is None else val          # a dip at `at` + 1, a boundary far outside the blob, a boundary below the chunk's start
        with pytest.raises(engine.ClkError) as ei:
            e.classify_varlen(blob, bad, abi.F_NAME_STAGE)
        assert ei.value.code == abi.EINVAL and "monotonic" in str(ei.value)
    assert np.array_equal(e.classify_varlen(blob, offs, abi.F_NAME_STAGE), good)      # the engine is still usable
    e.close()
