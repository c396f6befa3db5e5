"""Multi-GPU plumbing for the one place the path shards: packet batches are split into
contiguous index ranges, tables are replicated, and the per-GPU metrics_map tables are summed across GPUs —
the analogue of summing the reference's BPF_MAP_TYPE_PERCPU_HASH values (internal/ebpf/bpf/common.h:171-177).
Two forms of that sum: after every batch a small dense block of per-container ALLOW/DENY/BYPASS totals is
all-gathered (32 B per container, cost independent of the key count and of the number of GPUs that matter);
the per-key merge (all-gather of the compacted tables) runs when somebody reads the merged map.
One process per GPU, torch.distributed for the collectives."""
from __future__ import annotations

import numpy as np

from .abi import METRIC_DTYPE, TOTAL_DTYPE


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """GPU `rank` classifies packets [n*rank/world, n*(rank+1)/world)."""
    return n * rank // world, n * (rank + 1) // world


def rows_to_block(rows: np.ndarray, block_rows: int) -> np.ndarray:
    """Fixed-size counter block (count == 0 rows are empty), the payload shape of the all-gather."""
    if rows.size > block_rows:
        raise ValueError(f"{rows.size} metric rows do not fit a block of {block_rows}")
    blk = np.zeros(block_rows, dtype=METRIC_DTYPE)
    blk[: rows.size] = rows
    return blk


def merge_blocks(blocks: np.ndarray) -> np.ndarray:
    """Sum gathered blocks by metric key (host statement of k_counters_import)."""
    rows = blocks.reshape(-1)
    rows = rows[rows["count"] > 0]
    if rows.size == 0:
        return rows
    keys = np.stack([rows["cgroup_id"], rows["domain_hash"].astype(np.uint64), rows["dst_port"].astype(np.uint64),
                     rows["action"].astype(np.uint64)], axis=1)
    uniq, inv = np.unique(keys, axis=0, return_inverse=True)
    out = np.zeros(uniq.shape[0], dtype=METRIC_DTYPE)
    out["cgroup_id"], out["domain_hash"], out["dst_port"], out["action"] = uniq[:, 0], uniq[:, 1], uniq[:, 2], uniq[:, 3]
    np.add.at(out["count"], inv.reshape(-1), rows["count"])
    return out


def allgather_counters(engine, group=None) -> None:
    """The per-batch collective on the GPU path: export this GPU's table, ncclAllGather, sum into the
    engine's merged view (Engine.counters(merged=True)).  Runs on a stream of its own (the engine's kernels and
    the NCCL kernel must be ordered on ONE stream; torch's default stream has handle 0, which the C ABI reads as
    "the engine's stream") after everything already enqueued on the device; returns when the merge is done."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rows = engine.counters_block_rows()
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.cuda.synchronize()
    s = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(s):
        block = torch.empty(rows * METRIC_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        gathered = torch.empty(world * rows * METRIC_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        engine.counters_export(block.data_ptr(), s.cuda_stream)
        dist.all_gather_into_tensor(gathered, block, group=group)
        engine.counters_import(gathered.data_ptr(), world, s.cuda_stream)
    s.synchronize()


def totals_of(rows: np.ndarray, block_rows: int) -> np.ndarray:
    """Dense per-container totals block of one table (host statement of k_counters_totals; rows sorted by cgroup)."""
    cgs = np.unique(rows["cgroup_id"])
    if cgs.size > block_rows:
        raise ValueError(f"{cgs.size} containers do not fit a block of {block_rows}")
    blk = np.zeros(block_rows, dtype=TOTAL_DTYPE)
    blk["cgroup_id"][: cgs.size] = cgs
    idx = np.searchsorted(cgs, rows["cgroup_id"])
    np.add.at(blk["count"], (idx, rows["action"].astype(np.int64)), rows["count"])
    return blk


def merge_totals(blocks: np.ndarray) -> np.ndarray:
    """Sum gathered totals blocks by cgroup_id (host statement of clk_totals_merge)."""
    rows = blocks.reshape(-1)
    rows = rows[(rows["cgroup_id"] != 0) | (rows["count"].sum(axis=1) != 0)]
    cgs, inv = np.unique(rows["cgroup_id"], return_inverse=True)
    out = np.zeros(cgs.size, dtype=TOTAL_DTYPE)
    out["cgroup_id"] = cgs
    np.add.at(out["count"], inv.reshape(-1), rows["count"])
    return out


class TotalsGather:
    """Per-batch collective on the GPU path: k_counters_totals -> ncclAllGather of rows x 32 B.  Buffers are
    allocated once; `enqueue` only launches (no host synchronisation), `read` merges the last gathered blocks."""

    def __init__(self, engine, group=None):
        import torch
        import torch.distributed as dist
        self.e, self.group, self.world = engine, group, dist.get_world_size(group)
        self.rows = engine.totals_rows()
        dev = torch.device("cuda", torch.cuda.current_device())
        self.block = torch.zeros(self.rows * TOTAL_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        self.gathered = torch.zeros(self.world * self.rows * TOTAL_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        self.stream = torch.cuda.Stream(device=dev)              # side stream: overlaps the next batch's classify kernel
        torch.cuda.synchronize()

    def enqueue(self, after=None):
        """Launch export + all-gather on the side stream, ordered after `after` (a torch stream) if given."""
        import torch
        import torch.distributed as dist
        if after is not None:
            self.stream.wait_stream(after)
        with torch.cuda.stream(self.stream):
            self.e.totals_export(self.block.data_ptr(), self.stream.cuda_stream)
            dist.all_gather_into_tensor(self.gathered, self.block, group=self.group)

    def read(self) -> np.ndarray:
        return self.e.totals_merge(self.gathered.data_ptr(), self.world)
