"""Two-GPU form of the sharding test (runs only where >= 2 CUDA devices are visible, e.g. gpurun --gpus 2):
packet shards on two B200s, ncclAllGather of the counter blocks, merged view == oracle on the whole batch."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N = 400_003


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from clawker_b200 import capture, sharding
    from util import make_engine
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    w = capture.make_workload("c2")
    pk = capture.gen_slots(w, N, seed=4)
    lo, hi = sharding.shard_range(N, rank, world)
    e = make_engine(w, device=rank, max_metrics=1 << 18)
    verd = e.classify_slots(pk[lo:hi])
    tg = sharding.TotalsGather(e)
    tg.enqueue()                                              # per-batch form: dense per-container totals
    totals = tg.read()
    sharding.allgather_counters(e)                            # on-read form: per-key merge
    merged_torch = e.counters(merged=True)
    # the same two collectives behind the C ABI (clk_comm_init / clk_allgather_counters / clk_allgather_metrics): what a Go host calls
    from clawker_b200 import engine as eng
    ids = [eng.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    e.comm_init(ids[0], rank, world)
    e.allgather_counters()
    totals_abi = e.totals_global()
    e.allgather_metrics()
    merged_abi = e.counters(merged=True)
    e.comm_destroy()
    q.put((rank, lo, hi, verd, merged_torch, e.counters(), totals, totals_abi, merged_abi))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_shards_and_nccl_counter_allgather():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from clawker_b200 import capture, sharding
    from oracle import oracle
    from util import assert_same_metrics, oracle_tables
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w = capture.make_workload("c2")
    pk = capture.gen_slots(w, N, seed=4)
    want_v, want_m = oracle.Oracle("port").load(oracle_tables(w)).classify(pk)
    assert np.array_equal(np.concatenate([r[3] for r in res]), want_v)
    want_t = sharding.merge_totals(sharding.totals_of(want_m, 256))
    for rank, lo, hi, verd, merged, local, totals, totals_abi, merged_abi in res:
        assert_same_metrics(merged, want_m, f"rank {rank} merged")
        assert_same_metrics(merged_abi, want_m, f"rank {rank} merged through clk_allgather_metrics")
        totals_abi = totals_abi[totals_abi["count"].sum(axis=1) > 0]
        assert np.array_equal(totals_abi, want_t), f"rank {rank} totals through clk_allgather_counters"
        totals = totals[totals["count"].sum(axis=1) > 0]                     # containers without traffic have all-zero rows
        assert np.array_equal(totals, want_t), f"rank {rank} totals"
        assert int(local["count"].sum()) < int(merged["count"].sum())
