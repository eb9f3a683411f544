"""Region sharding across the GPUs of one box (one process per GPU).

The reference parallelises this path as an embarrassingly parallel map over regions
(`cmclapply(1:length(mask), coverageFromRanges, ...)`, /root/reference/R/coverage.R:148-154) and
reassembles the matrix with `do.call(rbind, ...)` (R/profile.R:150,208).  Here each rank owns a
slice of the regions (plus the reads that can overlap it), computes its row block on its own
GPU with no data-path collective, and the row blocks are gathered once at the end.

`torch.distributed` is plumbing only: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
import numpy as np


def partition_regions(chrom, start, end, world):
    """Split regions into `world` slices that are contiguous in (chrom, start) order and balanced
    by covered bases (gene lengths are heavy-tailed, so balancing by count is not enough).
    Returns a list of index arrays (into the caller's region order), one per rank."""
    chrom = np.asarray(chrom, dtype=np.int64)
    start = np.asarray(start, dtype=np.int64)
    end = np.asarray(end, dtype=np.int64)
    n = start.shape[0]
    order = np.lexsort((start, chrom))
    width = np.maximum(end[order] - start[order] + 1, 1)
    cum = np.cumsum(width)
    total = int(cum[-1]) if n else 0
    out = []
    lo = 0
    for r in range(world):
        target = total * (r + 1) / world
        hi = int(np.searchsorted(cum, target, side="left")) + 1 if r < world - 1 else n
        hi = min(max(hi, lo), n)
        out.append(order[lo:hi])
        lo = hi
    return out


def reads_for_slice(read_chrom, read_start, read_end, reg_chrom, reg_start, reg_end, frag_len=0):
    """Boolean mask of the reads that can overlap a region slice: per chromosome, reads inside
    [min region start - frag_len, max region end + frag_len].  Conservative (never drops an
    overlapping read); reads at slice boundaries are simply kept by both neighbours."""
    read_chrom = np.asarray(read_chrom)
    keep = np.zeros(read_chrom.shape[0], dtype=bool)
    reg_chrom = np.asarray(reg_chrom)
    for c in np.unique(reg_chrom):
        sel = reg_chrom == c
        lo = int(np.asarray(reg_start)[sel].min()) - int(frag_len)
        hi = int(np.asarray(reg_end)[sel].max()) + int(frag_len)
        keep |= (read_chrom == c) & (np.asarray(read_end) >= lo) & (np.asarray(read_start) <= hi)
    return keep


def gather_rows(local, row_ids, n_total, dst=0, group=None, scatter=None, sizes=None):
    """Gather per-rank row blocks into the full matrix on rank `dst`.

    local    torch tensor [n_cols, n_local]: a column-major  n_local x n_cols  block
    row_ids  int64 numpy array (n_local): the global row of each local row
    returns  on `dst` a torch tensor [n_cols, n_total] (column-major n_total x n_cols), else None
    scatter  optional callable(block, ids, k, full) placing the first k rows of a block on the
             device (bench.py passes rcp_rows_scatter); default is torch index_copy_.
    sizes    optional list of every rank's n_local (skips the object all-gather)
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_cols, n_local = int(local.shape[0]), int(local.shape[1])
    if sizes is None:
        sizes = [None] * world
        dist.all_gather_object(sizes, n_local, group=group)
    n_max = max(sizes) if sizes else 0
    ids = torch.full((n_max,), -1, dtype=torch.int64, device=local.device)
    ids[:n_local] = torch.as_tensor(np.asarray(row_ids, dtype=np.int64), device=local.device)
    block = torch.zeros((n_cols, n_max), dtype=local.dtype, device=local.device)
    block[:, :n_local] = local
    if rank == dst:
        blocks = [torch.empty_like(block) for _ in range(world)]
        id_list = [torch.empty_like(ids) for _ in range(world)]
    else:
        blocks = id_list = None
    dist.gather(block, blocks, dst=dst, group=group)
    dist.gather(ids, id_list, dst=dst, group=group)
    if rank != dst:
        return None
    full = torch.zeros((n_cols, n_total), dtype=local.dtype, device=local.device)
    for r in range(world):
        k = sizes[r]
        if k == 0:
            continue
        if scatter is not None:
            scatter(blocks[r], id_list[r][:k], k, full)
        else:
            full.index_copy_(1, id_list[r][:k], blocks[r][:, :k])
    return full
