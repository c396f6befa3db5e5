"""Multi-GPU plumbing for the one place the path shards: packet batches are split into
contiguous index ranges, tables are replicated, and the per-GPU metrics_map tables are all-gathered once
per batch and summed — the GPU analogue of summing the reference's BPF_MAP_TYPE_PERCPU_HASH values
(internal/ebpf/bpf/common.h:171-177).  One process per GPU, torch.distributed for the collective."""
from __future__ import annotations

import numpy as np

from .abi import METRIC_DTYPE


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
    engine's merged view (Engine.counters(merged=True)).  Enqueued on torch's current stream."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rows = engine.counters_block_rows()
    dev = torch.device("cuda", torch.cuda.current_device())
    block = torch.empty(rows * METRIC_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    gathered = torch.empty(world * rows * METRIC_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    engine.counters_export(block.data_ptr(), st)
    dist.all_gather_into_tensor(gathered, block, group=group)
    engine.counters_import(gathered.data_ptr(), world, st)
    torch.cuda.current_stream().synchronize()
