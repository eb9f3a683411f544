"""The N > 1 host path on CPU: world_size = 2 over gloo.  Each rank owns a region slice and the
reads that can overlap it, computes its row block (here with the ORACLE standing in for the GPU,
which CPU tests may do), and the blocks are gathered; the result must equal the single-process
matrix bit for bit."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    import workloads as W
    from oracle import c_oracle as CO
    from oracle import recoup_oracle as O
    from recoup_b200.sharding import RowGather, gather_rows, partition_regions, reads_for_slice

    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = W.gene_bodies(scale=0.004, seed=77)
    s, e = O.get_regional_ranges(w["region_start"], w["region_end"], w["region_strand"],
                                 "genebody", w["flank"])
    parts = partition_regions(w["region_chrom"], s, e, world)
    mine = parts[rank]
    keep = reads_for_slice(w["read_chrom"], w["read_start"], w["read_end"],
                           w["region_chrom"][mine], s[mine], e[mine])
    ix = CO.Index(w["read_chrom"][keep], w["read_start"][keep], w["read_end"][keep],
                  w["read_strand"][keep], w["chrom_len"])
    dense = CO.coverage(ix, w["region_chrom"][mine], s[mine], e[mine], w["region_strand"][mine])
    m = CO.profile_matrix(dense, w["flank"], w["bin_params"], False)          # [n_local x 250]
    local = torch.from_numpy(np.ascontiguousarray(m.T))                      # [250, n_local]
    full = gather_rows(local, mine, len(s), dst=0)
    # the persistent form: buffers and row indices set up once, then one gather per step
    g = RowGather(local.shape[0], mine, len(s), local.device, local.dtype, dst=0)
    for step in range(3):
        again = g.gather(local * float(step + 1))
        if rank == 0:
            assert np.array_equal(again.numpy(), full.numpy() * float(step + 1))
    if rank == 0:
        np.save(out_path, full.numpy().T)
    dist.barrier()
    dist.destroy_process_group()


def test_partition_is_a_balanced_permutation():
    from recoup_b200.sharding import partition_regions
    rng = np.random.default_rng(1)
    chrom = rng.integers(0, 5, size=1000)
    start = rng.integers(1, 10**6, size=1000)
    end = start + np.exp(rng.normal(8, 1.5, size=1000)).astype(np.int64)
    for world in (1, 2, 3, 8):
        parts = partition_regions(chrom, start, end, world)
        assert len(parts) == world
        allidx = np.concatenate(parts)
        assert sorted(allidx.tolist()) == list(range(1000))
        load = np.array([(end[p] - start[p] + 1).sum() for p in parts], dtype=float)
        assert load.max() <= load.mean() * 1.0 + (end - start + 1).max() * 2
        # contiguous in (chrom, start) order
        order = np.lexsort((start, chrom))
        assert np.array_equal(allidx, order)


def test_two_rank_gloo_matches_single_process(tmp_path):
    import torch.multiprocessing as mp

    import workloads as W
    from oracle import c_oracle as CO
    from oracle import recoup_oracle as O

    out = str(tmp_path / "full.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    w = W.gene_bodies(scale=0.004, seed=77)
    s, e = O.get_regional_ranges(w["region_start"], w["region_end"], w["region_strand"],
                                 "genebody", w["flank"])
    ix = CO.Index(w["read_chrom"], w["read_start"], w["read_end"], w["read_strand"], w["chrom_len"])
    dense = CO.coverage(ix, w["region_chrom"], s, e, w["region_strand"])
    want = CO.profile_matrix(dense, w["flank"], w["bin_params"], False)
    assert got.shape == want.shape == (len(s), 250)
    assert np.array_equal(got, want)        # sharded == single process, bitwise
