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


def slice_spans(reg_chrom, reg_start, reg_end, parts, n_chrom, pad=0):
    """Per rank and chromosome, the span [lo, hi] of the rank's region slice (lo > hi: the slice
    has no region on that chromosome), widened by `pad` (the fragment length when reads are
    extended on load).  int64 array [world, n_chrom, 2]; the input of exchange_reads."""
    reg_chrom = np.asarray(reg_chrom, dtype=np.int64)
    reg_start = np.asarray(reg_start, dtype=np.int64)
    reg_end = np.asarray(reg_end, dtype=np.int64)
    out = np.empty((len(parts), int(n_chrom), 2), dtype=np.int64)
    out[:, :, 0] = np.iinfo(np.int64).max // 4
    out[:, :, 1] = -1
    for r, idx in enumerate(parts):
        if len(idx) == 0:
            continue
        c = reg_chrom[idx]
        np.minimum.at(out[r, :, 0], c, reg_start[idx] - int(pad))
        np.maximum.at(out[r, :, 1], c, reg_end[idx] + int(pad))
    return out


def exchange_reads(chrom, start, end, strand, spans, group=None, filter_single=True):
    """The data-path exchange of the region-sharded run: every rank holds an arbitrary share of
    the reads (torch tensors on its device: chrom / start / end int32, strand int8 or None) and
    receives the reads that can overlap ITS region slice -- per chromosome, those meeting
    [lo, hi] of `spans` (slice_spans).  A read at a slice boundary goes to both neighbours.  One
    all_to_all_single for the counts, one for the packed (chrom, start, end) triples, one for the
    strands.  world == 1 (or no process group): a local filter against spans[0] (skipped with
    filter_single=False: the coverage call drops the reads its mask does not touch anyway).
    Returns the four tensors."""
    import torch
    import torch.distributed as dist

    on = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if on else 1
    if world == 1 and len(spans) == 1 and not filter_single:
        return chrom, start, end, strand       # one rank owns every region: nothing to send or drop
    dev = chrom.device
    if dev.type == "cuda":
        return _exchange_reads_device(chrom, start, end, strand, spans, world, group)
    # host tensors (the gloo tests of the N > 1 logic): the same routing with torch operations
    sp = torch.as_tensor(np.ascontiguousarray(spans), device=dev)        # [world, n_chrom, 2]
    c64 = chrom.long()
    picks = []
    for r in range(world):
        m = (end.long() >= sp[r, :, 0][c64]) & (start.long() <= sp[r, :, 1][c64])
        picks.append(torch.nonzero(m, as_tuple=False).squeeze(1))
    order = torch.cat(picks) if world > 1 else picks[0]
    triples = torch.stack([chrom[order], start[order], end[order]], dim=1).contiguous()
    st = strand[order].contiguous() if strand is not None else None
    if world == 1:
        return triples[:, 0].contiguous(), triples[:, 1].contiguous(), triples[:, 2].contiguous(), st
    send_counts = torch.tensor([int(p.shape[0]) for p in picks], dtype=torch.int64, device=dev)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    sc, rc = send_counts.tolist(), recv_counts.tolist()
    got = torch.empty((sum(rc), 3), dtype=triples.dtype, device=dev)
    dist.all_to_all_single(got, triples, output_split_sizes=rc, input_split_sizes=sc, group=group)
    got_st = None
    if st is not None:
        got_st = torch.empty((sum(rc),), dtype=st.dtype, device=dev)
        dist.all_to_all_single(got_st, st, output_split_sizes=rc, input_split_sizes=sc, group=group)
    return got[:, 0].contiguous(), got[:, 1].contiguous(), got[:, 2].contiguous(), got_st


def _exchange_reads_device(chrom, start, end, strand, spans, world, group):
    """exchange_reads on CUDA tensors: the library's routing kernels (rcp_reads_route_count /
    _pack) count and pack the reads per destination rank -- two streaming passes over the share,
    one host synchronisation -- then one all-to-all for the counts and one per array (no
    interleaving on the send side, nothing to take apart on the receive side).  The torch stream current at the call must be the library's stream."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    from . import _lib
    _lib.ensure_init()
    L = _lib.lib
    dev = chrom.device
    n = int(chrom.shape[0])
    sp = np.ascontiguousarray(np.clip(np.asarray(spans, dtype=np.int64), -2**31 + 1, 2**31 - 1), dtype=np.int32)
    n_chrom = int(sp.shape[1])
    chrom, start, end = chrom.contiguous(), start.contiguous(), end.contiguous()
    strand = None if strand is None else strand.contiguous()
    vp = lambda t: None if t is None else C.c_void_p(t.data_ptr())      # noqa: E731
    sp_p = sp.ctypes.data_as(C.POINTER(C.c_int32))
    counts = np.zeros(world, dtype=np.int64)
    _lib.check(L.rcp_reads_route_count(n, vp(chrom), vp(start), vp(end), world, n_chrom, sp_p,
                                       counts.ctypes.data_as(C.POINTER(C.c_int64))))
    offsets = np.ascontiguousarray(np.concatenate(([0], np.cumsum(counts)[:-1])), dtype=np.int64)
    total = int(counts.sum())
    send = [torch.empty((total,), dtype=torch.int32, device=dev) for _ in range(3)]
    st = torch.empty((total,), dtype=torch.int8, device=dev) if strand is not None else None
    _lib.check(L.rcp_reads_route_pack(n, vp(chrom), vp(start), vp(end), vp(strand), world, n_chrom, sp_p,
                                      offsets.ctypes.data_as(C.POINTER(C.c_int64)), vp(send[0]), vp(send[1]),
                                      vp(send[2]), vp(st)))
    if world == 1:
        return send[0], send[1], send[2], st
    send_counts = torch.from_numpy(counts).to(dev)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    sc, rc = counts.tolist(), recv_counts.tolist()
    got = []
    for t in send + ([st] if st is not None else []):
        r = torch.empty((sum(rc),), dtype=t.dtype, device=dev)
        dist.all_to_all_single(r, t, output_split_sizes=rc, input_split_sizes=sc, group=group)
        got.append(r)
    return got[0], got[1], got[2], (got[3] if st is not None else None)


class RowGather:
    """Gathers per-rank row blocks into the full matrix on rank `dst`, step after step: every
    buffer (padded send block, receive blocks, row-index lists, the full matrix) is allocated and
    the row indices are exchanged ONCE, so a step costs one NCCL gather plus the placement of the
    blocks.

    local    torch tensor [n_cols, n_local]: a column-major  n_local x n_cols  block
    row_ids  int64 numpy array (n_local): the global row of each local row
    scatter  optional callable(block, ids, k, full) placing the first k rows of a block on the
             device (bench.py passes rcp_rows_scatter); default is torch index_copy_.
    sizes    optional list of every rank's n_local (skips the object all-gather)
    """

    def __init__(self, n_cols, row_ids, n_total, device, dtype, dst=0, group=None, sizes=None):
        import torch
        import torch.distributed as dist

        self.group, self.dst = group, dst
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n_cols, self.n_total = int(n_cols), int(n_total)
        n_local = int(np.asarray(row_ids).shape[0])
        if sizes is None:
            sizes = [None] * self.world
            dist.all_gather_object(sizes, n_local, group=group)
        self.sizes = list(sizes)
        self.n_local = n_local
        self.n_max = max(self.sizes) if self.sizes else 0
        self.equal = all(k == self.n_max for k in self.sizes)
        ids = torch.full((self.n_max,), -1, dtype=torch.int64, device=device)
        ids[:n_local] = torch.as_tensor(np.asarray(row_ids, dtype=np.int64), device=device)
        self.block = None if self.equal else torch.zeros((self.n_cols, self.n_max), dtype=dtype,
                                                         device=device)
        if self.rank == dst:
            self.stack = torch.empty((self.world, self.n_cols, self.n_max), dtype=dtype, device=device)
            self.blocks = [self.stack[r] for r in range(self.world)]
            self.id_list = [torch.empty_like(ids) for _ in range(self.world)]
            self.full = torch.zeros((self.n_cols, self.n_total), dtype=dtype, device=device)
        else:
            self.stack = self.blocks = self.id_list = self.full = None
        dist.gather(ids, self.id_list, dst=dst, group=group)      # once
        # rank r owns rows [r * n, (r + 1) * n): the blocks drop into place with ONE strided copy
        self.in_order = False
        if self.rank == dst and self.equal and self.n_max * self.world == self.n_total:
            want = torch.arange(self.n_total, dtype=torch.int64, device=device)
            self.in_order = bool(torch.equal(torch.cat(self.id_list), want))

    def gather(self, local, scatter=None):
        """returns on `dst` a torch tensor [n_cols, n_total] (column-major n_total x n_cols, reused
        by the next call), else None"""
        import torch.distributed as dist

        if self.equal:
            send = local if local.is_contiguous() else local.contiguous()
        else:
            self.block[:, :self.n_local] = local
            send = self.block
        dist.gather(send, self.blocks, dst=self.dst, group=self.group)
        if self.rank != self.dst:
            return None
        if self.in_order and scatter is None:
            self.full.view(self.n_cols, self.world, self.n_max).copy_(self.stack.permute(1, 0, 2))
            return self.full
        for r in range(self.world):
            k = self.sizes[r]
            if k == 0:
                continue
            if scatter is not None:
                scatter(self.blocks[r], self.id_list[r][:k], k, self.full)
            else:
                self.full.index_copy_(1, self.id_list[r][:k], self.blocks[r][:, :k])
        return self.full


class PeerMatrix:
    """The full profile matrix on rank `dst`, mapped into every rank of the box (CUDA IPC through
    rcp_shared_alloc / rcp_shared_open): each rank's bin kernels store their row block straight
    into it over NVLink (`out = ptr_for(first_row)`, `ld = n_total`), so the exchange is fused
    into the kernel's own stores -- no gather, no scatter pass.  `fence()` (a tiny all-reduce on
    the caller's stream) orders every rank's stores before the matrix is read on `dst`.

    Contiguous row slices only (rank r owns rows [first_row, first_row + n_local)); arbitrary
    row sets use RowGather."""

    def __init__(self, n_total, n_cols, device, dst=0, group=None):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _lib

        self._lib = _lib
        self.group, self.dst = group, dst
        self.rank = dist.get_rank(group)
        self.n_total, self.n_cols = int(n_total), int(n_cols)
        nbytes = self.n_total * self.n_cols * 8
        handle = torch.zeros(_IPC_BYTES, dtype=torch.uint8)
        ptr = C.c_void_p(0)
        if self.rank == dst:
            buf = (C.c_ubyte * _IPC_BYTES)()
            _lib.check(_lib.lib.rcp_shared_alloc(nbytes, C.byref(ptr), buf))
            handle = torch.tensor(list(buf), dtype=torch.uint8)
        handle = handle.to(device)
        dist.broadcast(handle, src=dst, group=group)
        if self.rank != dst:
            buf = (C.c_ubyte * _IPC_BYTES)(*handle.cpu().tolist())
            _lib.check(_lib.lib.rcp_shared_open(buf, C.byref(ptr)))
        self.base = ptr.value
        self._flag = torch.zeros(1, dtype=torch.int32, device=device)
        dist.barrier(group=group)

    def ptr_for(self, first_row):
        """device pointer of element (first_row, 0): pass it as `out` with ld = n_total"""
        return self.base + 8 * int(first_row)

    def put(self, local, first_row, stream=None):
        """Copy-engine exchange: the block `local` ([n_cols, n_local] torch tensor = column-major
        n_local x n_cols) goes to rows [first_row, first_row + n_local) of the matrix on `dst` with
        one strided copy over NVLink (rcp_rows_put) on `stream` (a torch stream; default: the
        current one).  No SM of either GPU is used; follow with fence()."""
        import ctypes as C

        import torch
        st = stream if stream is not None else torch.cuda.current_stream(local.device)
        n_cols, n_local = int(local.shape[0]), int(local.shape[1])
        assert n_cols == self.n_cols and local.stride(1) == 1
        self._lib.check(self._lib.lib.rcp_rows_put(C.c_void_p(local.data_ptr()), int(local.stride(0)), n_local,
                                                   n_cols, C.c_void_p(self.ptr_for(first_row)), self.n_total,
                                                   C.c_void_p(st.cuda_stream)))

    def fence(self):
        """Orders every rank's stores / puts before the matrix is read on `dst` (a one-word
        all-reduce on the current torch stream)."""
        import torch.distributed as dist
        dist.all_reduce(self._flag, group=self.group)

    def as_tensor(self):
        """[n_cols, n_total] float64 view on rank `dst` (column-major n_total x n_cols)"""
        import torch
        assert self.rank == self.dst

        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (self.n_cols, self.n_total), "typestr": "<f8",
                                        "data": (self.base, False), "version": 2}
        return torch.as_tensor(raw, device=self._flag.device)

    def close(self):
        import torch.distributed as dist
        dist.barrier(group=self.group)
        if self.base:
            if self.rank == self.dst:
                self._lib.lib.rcp_shared_free(self.base)
            else:
                self._lib.lib.rcp_shared_close(self.base)
            self.base = 0


_IPC_BYTES = 64


def gather_rows(local, row_ids, n_total, dst=0, group=None, scatter=None, sizes=None):
    """One-shot form of RowGather (allocates per call)."""
    g = RowGather(int(local.shape[0]), row_ids, n_total, local.device, local.dtype, dst=dst,
                  group=group, sizes=sizes)
    return g.gather(local, scatter=scatter)
