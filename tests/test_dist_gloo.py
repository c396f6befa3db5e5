"""world_size-2 (gloo, CPU) test of the N>1 host logic: contiguous packet shards, replicated tables, one
all-gather of fixed-size counter blocks per batch, summed view == single-process result.  The per-shard
classification here is done by the oracle (no GPU in this container); the GPU form of the same test is
tests/test_gpu_multi.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from clawker_b200 import capture, sharding
from clawker_b200.abi import METRIC_DTYPE, TOTAL_DTYPE
from oracle import oracle

from util import assert_same_metrics, oracle_tables

N, BLOCK_ROWS = 40_001, 16384


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = capture.make_workload("c2")
    pk = capture.gen_slots(w, N, seed=4)                      # every rank holds the same capture, classifies its shard
    lo, hi = sharding.shard_range(N, rank, world)
    verd, met = oracle.Oracle("port").load(oracle_tables(w)).classify(pk[lo:hi])
    blk = torch.from_numpy(sharding.rows_to_block(met, BLOCK_ROWS).view(np.uint8).copy())
    out = [torch.empty_like(blk) for _ in range(world)]
    dist.all_gather(out, blk)                                 # the one collective
    merged = sharding.merge_blocks(np.stack([o.numpy().view(METRIC_DTYPE) for o in out]))
    tb = torch.from_numpy(sharding.totals_of(met, 16).view(np.uint8).copy())          # the per-batch totals block
    tout = [torch.empty_like(tb) for _ in range(world)]
    dist.all_gather(tout, tb)
    totals = sharding.merge_totals(np.stack([o.numpy().view(TOTAL_DTYPE) for o in tout]))
    vt = [None] * world
    dist.all_gather_object(vt, (lo, hi, verd))
    q.put((rank, merged, vt, totals))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_rank_shards_and_counter_allgather(world):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w = capture.make_workload("c2")
    pk = capture.gen_slots(w, N, seed=4)
    want_v, want_m = oracle.Oracle("port").load(oracle_tables(w)).classify(pk)
    want_t = sharding.merge_totals(sharding.totals_of(want_m, 16))
    for rank, merged, vt, totals in res:
        assert_same_metrics(merged, want_m, f"rank {rank} merged view")          # every rank ends with the same sum
        assert np.array_equal(totals, want_t) and int(totals["count"].sum()) == int(want_m["count"].sum())
        got = np.concatenate([v for _, _, v in sorted(vt, key=lambda t: t[0])])
        assert np.array_equal(got, want_v)                                        # verdicts independent of the sharding


def test_shard_ranges_cover_exactly():
    for n in (0, 1, 7, 1000, 10_000_019):
        for world in (1, 2, 4, 8):
            r = [sharding.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))


def test_merge_blocks_sums_by_key():
    a = np.zeros(3, dtype=METRIC_DTYPE)
    a[0] = (1, 2, 3, 0, 0, 5)
    a[1] = (1, 2, 3, 1, 0, 7)
    b = np.zeros(3, dtype=METRIC_DTYPE)
    b[0] = (1, 2, 3, 0, 0, 11)
    m = sharding.merge_blocks(np.stack([a, b]))
    assert m.size == 2 and int(m["count"].sum()) == 23
    with pytest.raises(ValueError):
        sharding.rows_to_block(np.zeros(5, dtype=METRIC_DTYPE), 4)


def test_totals_blocks():
    a = np.zeros(4, dtype=METRIC_DTYPE)
    a[0] = (7, 2, 3, 0, 0, 5)
    a[1] = (7, 9, 3, 1, 0, 7)
    a[2] = (9, 2, 3, 2, 0, 1)
    a[3] = (7, 4, 4, 0, 0, 2)
    t = sharding.totals_of(a, 8)
    assert t["cgroup_id"][:2].tolist() == [7, 9] and t["count"][0].tolist() == [7, 7, 0] and t["count"][1].tolist() == [0, 0, 1]
    m = sharding.merge_totals(np.stack([t, t]))
    assert m.size == 2 and m["count"][0].tolist() == [14, 14, 0]
    with pytest.raises(ValueError):
        sharding.totals_of(a, 1)
