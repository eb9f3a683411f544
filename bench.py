#!/usr/bin/env python
"""bench.py -- reads -> coverage -> profile matrix on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2|C3|C4|C5] [--scale S]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      (CPU arm: the oracle's C port on the host cores)

A "step" is one pass of the hot path over one sample: rcp_reads_load (global-coordinate map +
index sort) -> rcp_coverage (exact per-base int32 coverage of every region) ->
rcp_profile_matrix (regions x bins fp64, column-major) [-> NCCL gather of the row blocks to rank 0
when N > 1].  `value` times that with the decoded reads/regions already resident in HBM;
`e2e` times the same through the public host API with pinned HOST buffers, copies included.
Regions shard across ranks with no data-path collective except the final row gather (weak
scaling: every rank owns a full-size region slice and its overlapping reads).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import workloads as W  # noqa: E402

METRIC = "reads/s through coverage + profileMatrix (reads -> per-base coverage -> regions x bins)"
UNIT = "reads/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(W.CONFIGS))
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fence", default="step", choices=["step", "end"],
                    help="p2p exchange: all-reduce fence after every step, or only once after the "
                         "timed steps (diagnostic: shows what the per-step fence costs)")
    ap.add_argument("--exchange", default="nccl", choices=["p2p", "nccl"],
                    help="N > 1: ranks store their rows straight into rank 0's matrix over NVLink "
                         "(p2p) or the row blocks are gathered with NCCL and placed by a kernel (nccl)")
    ap.add_argument("--read-order", default="random", choices=["random", "coordinate"],
                    help="order of the synthetic reads: as generated (random; the default and the "
                         "harder case) or sorted by chromosome and start like a coordinate-sorted BAM")
    ap.add_argument("--path", default="auto", choices=["auto", "index", "buckets", "blocks", "split"],
                    help="rcp_set_coverage_path: how rcp_coverage finds each region's reads")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons, sampled every 50 ms from before the warm-up until
    after the e2e loop; rows are stamped on arrival so the timed window can be cut out."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 5.0:
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t_begin, t_end, label="timed region"):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def digest(rows):
            sm, mx, reasons = [], [], set()
            for _t, row in rows:
                f = [x.strip() for x in row.split(",")]
                if len(f) < 6:
                    continue
                try:
                    sm.append(float(f[0]))
                    mx.append(float(f[1]))
                except ValueError:
                    continue
                for nme, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            return sm, mx, reasons

        inside = [r for r in self.rows if t_begin <= r[0] <= t_end]
        window = label
        if len(inside) < 3:
            inside = self.rows          # timed region shorter than the sampling period
            window = "warm-up + timed region (timed region < 3 samples long)"
        sm, mx, reasons = digest(inside)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port)
# ------------------------------------------------------------------------------------------------
def cpu_pass(w, frac):
    """One pass of the oracle's C port over the first `frac` of the reads and regions of
    workload `w`.  Returns (seconds, reads processed, description)."""
    from oracle import c_oracle as CO
    from oracle import recoup_oracle as O
    n = max(int(len(w["read_start"]) * frac), 1)
    t0 = time.perf_counter()
    ix = CO.Index(w["read_chrom"][:n], w["read_start"][:n], w["read_end"][:n], w["read_strand"][:n],
                  w["chrom_len"], frag_len=w["frag_len"])
    if w["region"] == "rna":
        G = max(int((len(w["exon_ptr"]) - 1) * frac), 1)
        ptr = w["exon_ptr"][:G + 1]
        ne = int(ptr[-1])
        f1, f2 = w["flank"]
        gs, ge, gst, gc = (w["region_start"][:G], w["region_end"][:G], w["region_strand"][:G],
                           w["region_chrom"][:G])
        ls, le = O.get_flanking_ranges(gs, ge, gst, f1, "upstream")
        rs, re_ = O.get_flanking_ranges(gs, ge, gst, f2, "downstream")
        center = CO.coverage_list(ix, ptr, w["exon_chrom"][:ne], w["exon_start"][:ne],
                                  w["exon_end"][:ne], w["exon_strand"][:ne])
        left = CO.coverage(ix, gc, ls, le, gst)
        right = CO.coverage(ix, gc, rs, re_, gst)
        # merge + bins on the list form (python); small next to the coverage passes
        merged = CO.concat3(left.to_list(), center.to_list(), right.to_list())
        O.profile_matrix(merged, w["flank"], w["bin_params"])
        n_units = G
    else:
        R = max(int(len(w["region_start"]) * frac), 1)
        s, e = O.get_regional_ranges(w["region_start"][:R], w["region_end"][:R],
                                     w["region_strand"][:R], w["region"], w["flank"])
        dense = CO.coverage(ix, w["region_chrom"][:R], s, e, w["region_strand"][:R])
        lens = dense.len[dense.len > 0]
        equal = bool((lens == lens[0]).all()) if lens.size else True
        CO.profile_matrix(dense, w["flank"], w["bin_params"], equal)
        n_units = R
    dt = time.perf_counter() - t0
    return dt, n, "first %d reads and %d regions of %s" % (n, n_units, w["name"])


def cpu_baseline(w, budget_s=20.0):
    from oracle import c_oracle as CO
    frac = min(1.0, 2_000_000 / max(len(w["read_start"]), 1))
    dt, n, what = cpu_pass(w, frac)                       # calibration pass
    rate = n / dt
    frac2 = min(1.0, budget_s * rate / len(w["read_start"]))
    if frac2 > frac * 1.5:
        dt, n, what = cpu_pass(w, frac2)
    return {"value": n / dt, "unit": UNIT, "cores": CO.threads(), "kind": "port",
            "sample": what + " (oracle C port, OpenMP over regions; R itself is not installable "
                             "here)", "seconds": round(dt, 3)}


def run_reference(args):
    """--impl reference: the reference's CPU path.  R/Bioconductor cannot run in this image (no
    R, no network), so this arm times the oracle's C/OpenMP port with every host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = W.CONFIGS[args.workload](scale=args.scale)
    from oracle import c_oracle as CO
    total = max(len(w["read_start"]), 1)
    frac = min(1.0, 2_000_000 / total)
    dt, n, what = cpu_pass(w, frac)                       # calibration
    per_step_s = min(20.0, 150.0 / max(args.steps + args.warmup, 1))
    frac = min(1.0, max(frac, per_step_s * (n / dt) / total))
    for _ in range(args.warmup):
        cpu_pass(w, frac)
    t = 0.0
    reads = 0
    for _ in range(args.steps):
        dt, n, what = cpu_pass(w, frac)
        t += dt
        reads += n
    val = reads / t
    cb = {"value": val, "unit": UNIT, "cores": CO.threads(), "kind": "port", "sample": what}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/f64",
        "data": "synthetic", "config": {"workload": w["name"], "sample": what},
        "cpu_baseline": cb,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    import recoup_b200 as rb
    from recoup_b200 import _lib
    from recoup_b200.ranges import getFlankingRanges, getRegionalRanges

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # a mismatched collective must fail in a minute, not hang the box until the watchdog
        dist.init_process_group("nccl", device_id=torch.device("cuda", local),
                                timeout=datetime.timedelta(seconds=90))
    rb.init(local)
    rb.set_coverage_path(args.path)
    L = _lib.lib
    dev = torch.device("cuda", local)
    stream = torch.cuda.ExternalStream(L.rcp_stream(), device=dev)

    # ---- this rank's slice: a full-size sample of the named workload (weak scaling) ----
    w = W.CONFIGS[args.workload](scale=args.scale, seed={"C2": 1001, "C3": 1003, "C4": 1004,
                                                          "C5": 1005}[args.workload] + 7919 * rank)
    N = len(w["read_start"])
    if args.read_order == "coordinate":
        order = np.lexsort((w["read_start"], w["read_chrom"]))
        for key in ("read_chrom", "read_start", "read_end", "read_strand"):
            w[key] = np.ascontiguousarray(w[key][order])
        del order
        w["name"] += " (reads coordinate-sorted)"
    is_rna = w["region"] == "rna"
    genes = rb.GRanges(w["region_chrom"], w["region_start"], w["region_end"],
                       strand=w["region_strand"], seqlevels=w["chrom_names"])
    f1, f2 = w["flank"]
    bp = w["bin_params"]
    if is_rna:
        left = getFlankingRanges(genes, f1, "upstream")
        right = getFlankingRanges(genes, f2, "downstream")
        R = len(genes)
    else:
        win = getRegionalRanges(genes, w["region"], w["flank"])
        R = len(win)
    clen = np.ascontiguousarray(w["chrom_len"], dtype=np.int64)
    clen_p = clen.ctypes.data_as(C.POINTER(C.c_int64))
    n_chrom = clen.shape[0]

    def to_dev(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    d_reads = [to_dev(w[k]) for k in ("read_chrom", "read_start", "read_end", "read_strand")]
    if is_rna:
        d_left = [to_dev(x) for x in (left.seqnames, left.start, left.end, left.strand)]
        d_right = [to_dev(x) for x in (right.seqnames, right.start, right.end, right.strand)]
        ex = [np.ascontiguousarray(w[k]) for k in ("exon_chrom", "exon_start", "exon_end", "exon_strand")]
        ex_ptr = np.ascontiguousarray(w["exon_ptr"], dtype=np.int64)
    else:
        d_win = [to_dev(x) for x in (win.seqnames, win.start, win.end, win.strand)]
    vp = lambda t: C.c_void_p(t.data_ptr())
    hp = lambda a: a.ctypes.data_as(C.c_void_p)

    equal_lengths = 1 if w["region"] in ("tss", "tes", "custom") else 0
    ncols_box = {}
    out_box = {}
    stats = {}

    def device_step(k=0):
        """reads (HBM) -> index -> coverage -> matrix (HBM)."""
        h = C.c_int(0)
        _lib.check(L.rcp_reads_load(N, vp(d_reads[0]), vp(d_reads[1]), vp(d_reads[2]), vp(d_reads[3]),
                                    n_chrom, clen_p, int(w["frag_len"]), _lib.MEM_DEVICE, C.byref(h)))
        cov = C.c_int(0)
        if is_rna:
            hc, hl, hr = C.c_int(0), C.c_int(0), C.c_int(0)
            _lib.check(L.rcp_coverage_list(h.value, R, ex_ptr.ctypes.data_as(C.POINTER(C.c_int64)),
                                           hp(ex[0]), hp(ex[1]), hp(ex[2]), hp(ex[3]), 1,
                                           _lib.STRAND_ANY, _lib.MEM_HOST, C.byref(hc)))
            _lib.check(L.rcp_coverage(h.value, R, vp(d_left[0]), vp(d_left[1]), vp(d_left[2]),
                                      vp(d_left[3]), 1, _lib.STRAND_ANY, _lib.MEM_DEVICE, C.byref(hl)))
            _lib.check(L.rcp_coverage(h.value, R, vp(d_right[0]), vp(d_right[1]), vp(d_right[2]),
                                      vp(d_right[3]), 1, _lib.STRAND_ANY, _lib.MEM_DEVICE, C.byref(hr)))
            _lib.check(L.rcp_coverage_concat3(hl.value, hc.value, hr.value, C.byref(cov)))
            for x in (hc, hl, hr):
                L.rcp_coverage_free(x.value)
        else:
            _lib.check(L.rcp_coverage(h.value, R, vp(d_win[0]), vp(d_win[1]), vp(d_win[2]),
                                      vp(d_win[3]), 1, _lib.STRAND_ANY, _lib.MEM_DEVICE, C.byref(cov)))
        if "n" not in ncols_box:
            nc = C.c_int64(0)
            _lib.check(L.rcp_profile_ncols(cov.value, equal_lengths, f1, f2, bp["flankBinSize"],
                                           bp["regionBinSize"], C.byref(nc)))
            ncols_box["n"] = nc.value
            # col-major R x nc, double buffered (N > 1: the gather of step k overlaps step k + 1)
            out_box["bufs"] = [torch.empty((nc.value, R), dtype=torch.float64, device=dev)
                               for _ in range(2 if world > 1 else 1)]
            out_box["m"] = out_box["bufs"][0]
            out_box["ptr"], out_box["ld"] = None, R
            if world > 1 and args.exchange == "p2p":
                # every rank writes its rows straight into rank 0's matrix (peer-mapped)
                from recoup_b200.sharding import PeerMatrix
                with torch.cuda.stream(stream):
                    pm = PeerMatrix(R * world, nc.value, dev, dst=0)
                out_box["peer"] = pm
                out_box["ptr"], out_box["ld"] = pm.ptr_for(rank * R), R * world
            tl, nn = C.c_int64(0), C.c_int64(0)
            L.rcp_coverage_info(cov.value, None, C.byref(tl), C.byref(nn), None)
            stats["total_len"], stats["n_null"] = tl.value, nn.value
            pth, cand = C.c_int(0), C.c_int64(0)
            L.rcp_coverage_path_info(cov.value, C.byref(pth), C.byref(cand))
            stats["path"] = {0: "list", 1: "index", 2: "buckets", 3: "blocks", 4: "split"}[pth.value]
            stats["candidates"] = cand.value
        if out_box["ptr"] is not None:      # peer-mapped matrix of rank 0 / verification buffer
            out_ptr = out_box["ptr"]
        else:
            buf = out_box["bufs"][k % len(out_box["bufs"])]
            if world > 1:
                gissued[k % 2].wait()       # the helper thread has issued that gather
            if world > 1 and gdone[k % 2] is not None:
                stream.wait_event(gdone[k % 2])     # the gather that read this buffer two steps ago
            out_ptr = buf.data_ptr()
        _lib.check(L.rcp_profile_matrix(cov.value, equal_lengths, f1, f2, bp["flankBinSize"],
                                        bp["regionBinSize"], _lib.STAT[bp["sumStat"]],
                                        _lib.INTERP[bp["interpolation"]], 42, 0,
                                        C.c_void_p(out_ptr), out_box["ld"], _lib.MEM_DEVICE))
        L.rcp_coverage_free(cov.value)
        L.rcp_reads_free(h.value)

    gather_box = {}

    gstream = torch.cuda.Stream(device=dev) if world > 1 else None
    gdone = [None, None]

    def gather_issue(k, ready):
        from recoup_b200.sharding import RowGather
        gstream.wait_event(ready)
        with torch.cuda.stream(gstream):
            if "g" not in gather_box:       # buffers and row indices are set up once
                gather_box["g"] = RowGather(out_box["bufs"][0].shape[0],
                                            np.arange(rank * R, (rank + 1) * R), R * world, dev,
                                            out_box["bufs"][0].dtype, dst=0, sizes=[R] * world)
            gather_box["full"] = gather_box["g"].gather(out_box["bufs"][k % 2])
            done = torch.cuda.Event()
            done.record(gstream)
        return done

    # The host is on the critical path of a step (the library synchronises to validate the reads
    # and to size the hit list), so the NCCL call is issued by a helper thread while the main
    # thread is already launching the next step; ctypes releases the GIL inside the library.
    import queue
    import threading
    gq = queue.Queue()
    gissued = [threading.Event(), threading.Event()]
    for e in gissued:
        e.set()

    gfail = []

    def gather_worker():
        torch.cuda.set_device(dev)
        while True:
            item = gq.get()
            if item is None:
                return
            k, ready = item
            try:
                gdone[k % 2] = gather_issue(k, ready)
            except BaseException as exc:    # never leave the main thread waiting for this gather
                gfail.append(exc)
            finally:
                gissued[k % 2].set()

    gthread = None
    if world > 1 and args.exchange == "nccl" and not os.environ.get("RCP_BENCH_INLINE_GATHER"):
        gthread = threading.Thread(target=gather_worker, daemon=True)
        gthread.start()

    def gather_drain():
        for e in gissued:
            e.wait()
        if gfail:
            raise gfail[0]

    def gather_step(k=0):
        """NCCL gather of the row blocks to rank 0 (the reference's do.call(rbind, ...)) and their
        placement into the full column-major matrix.  It runs on its own stream, behind the bin
        kernel of this step and beside the next step's kernels: the output matrix is double
        buffered (step k writes buffer k % 2), so only the gather of step k - 2 must have finished
        before step k may overwrite its buffer."""
        if world == 1 or os.environ.get("RCP_BENCH_NO_GATHER"):
            return
        if "peer" in out_box:               # the rows are already in place: order the stores
            if args.fence == "step":
                with torch.cuda.stream(stream):
                    out_box["peer"].fence()
            return
        ready = torch.cuda.Event()
        ready.record(stream)                # the bin kernel of this step (library stream)
        if gthread is None:
            gdone[k % 2] = gather_issue(k, ready)
        else:
            gissued[k % 2].clear()
            gq.put((k, ready))

    def barrier():
        gather_drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    sampler = ClockSampler(local)
    sampler.start()
    for k in range(max(args.warmup, 3)):
        device_step(k)
        gather_step(k)
    barrier()

    # ---- timed: device-resident ----
    L.rcp_launch_count(1)
    _lib.check(L.rcp_timing_enable(1))
    _lib.check(L.rcp_timing_read(1, 0, None, None))
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.time()
    ev0.record(stream)
    for k in range(args.steps):
        device_step(k)
        gather_step(k)
    if world > 1:
        gather_drain()
        for ev in gdone:                    # the timed region ends when the last gathers have landed
            if ev is not None:
                stream.wait_event(ev)
    ev1.record(stream)
    barrier()
    t_end = time.time()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = int(L.rcp_launch_count(0))
    n_st = 0
    while L.rcp_timing_stage_name(n_st):
        n_st += 1
    ms = (C.c_double * n_st)()
    cnt = (C.c_int64 * n_st)()
    _lib.check(L.rcp_timing_read(1, n_st, ms, cnt))
    _lib.check(L.rcp_timing_enable(0))
    stage = {L.rcp_timing_stage_name(i).decode(): (ms[i] / max(cnt[i], 1), int(cnt[i]))
             for i in range(n_st) if cnt[i] > 0}
    # The timed region is a few tens of milliseconds, shorter than one nvidia-smi period: keep
    # the same steps running (untimed) for a second so that the clock record is taken UNDER THIS
    # LOAD.  The sampler polls nvidia-smi (a driver-lock heavy call), so it is stopped before the
    # host-API (e2e) loop.
    # The number of extra steps is agreed by all ranks (rank 0 derives it from the max-over-ranks
    # step time and broadcasts it): every rank issues the same number of gathers.
    n_probe = torch.tensor([max(1, min(4000, int(1000.0 / max(elapsed_ms / args.steps, 1e-3))))],
                           dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(n_probe, op=dist.ReduceOp.MIN)
    for k in range(args.steps, args.steps + int(n_probe.item())):
        device_step(k)
        gather_step(k)
    barrier()
    clocks = sampler.stop(t_begin, time.time(), "timed region + 1 s of the same steps (untimed)")

    # ---- timed: end to end through the public host API (pinned host buffers) ----
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a[:1]).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t

    pins = [pinned(w[k]) for k in ("read_chrom", "read_start", "read_end", "read_strand")]
    host_views = [p.numpy() for p in pins]
    h2d = sum(v.nbytes for v in host_views)
    if is_rna:
        grl = rb.GRangesList(rb.GRanges(w["exon_chrom"], w["exon_start"], w["exon_end"],
                                        strand=w["exon_strand"], seqlevels=w["chrom_names"]),
                             w["exon_ptr"])
        h2d += sum(ex[i].nbytes for i in range(4)) + 2 * 13 * R
    else:
        h2d += 13 * R
    ncols = ncols_box["n"]
    d2h = 8 * R * ncols + 4 * R
    e2e_steps = max(2, min(args.steps, 5))

    phases = {"upload+index": 0.0, "coverage": 0.0, "profile+download": 0.0}
    e2e_in = {"seqnames": host_views[0], "views": host_views}

    def e2e_step():
        """The calls a user of the reference API makes, on pinned HOST arrays."""
        t_a = time.perf_counter()
        hv = e2e_in["views"]
        reads = rb.GRanges(e2e_in["seqnames"], hv[1], hv[2], strand=hv[3],
                           seqlevels=w["chrom_names"], seqlengths=clen)
        sample = [dict(id="s", name="s", ranges=reads)]
        rb.device_reads(reads, w["frag_len"])
        L.rcp_sync()
        t_b = time.perf_counter()
        if is_rna:
            rb.coverageRnaRef(sample, grl, genes, w["flank"])
        else:
            if w["frag_len"]:
                sample[0]["coverage"] = rb.calcCoverage(reads, win, frag_len=w["frag_len"])
            else:
                rb.coverageRef(sample, genes, w["region"], w["flank"])
        L.rcp_sync()
        t_c = time.perf_counter()
        rb.profileMatrix(sample, w["flank"], bp)
        m = sample[0]["profile"]
        t_d = time.perf_counter()
        sample[0]["coverage"].free()
        for dr in reads._device.values():
            dr.free()
        reads._device.clear()
        phases["upload+index"] += t_b - t_a
        phases["coverage"] += t_c - t_b
        phases["profile+download"] += t_d - t_c
        return m

    # warm-up: two results alive at once, so that the library's pinned-buffer pool holds the two
    # output matrices the timed loop alternates between (page-locking 32 MB costs ~12 ms)
    keep = [e2e_step(), e2e_step()]
    del keep
    for k in phases:
        phases[k] = 0.0
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        mat = e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    assert mat.shape == (R, ncols)
    main_phases = dict(phases)

    # ---- the same, with the reads grouped by chromosome as a coordinate-sorted BAM delivers
    # them (random order inside a chromosome: nothing on the device assumes an order).  A GRanges
    # holds such seqnames as ~25 runs (Rle), and rcp_reads_load_rle sends the runs instead of
    # 4 bytes per read.  Reported beside `e2e`, which stays on the randomly ordered arrays.
    order = np.argsort(host_views[0].astype(np.uint8), kind="stable")
    for p in pins:
        p.numpy()[...] = p.numpy()[order]
    del order
    seq_rle = rb.Rle.encode(host_views[0])
    e2e_in["seqnames"] = seq_rle
    h2d_bam = h2d - host_views[0].nbytes + 8 * seq_rle.nrun
    want_sum = float(mat.sum())
    keep = [e2e_step(), e2e_step()]
    assert abs(float(keep[0].sum()) - want_sum) <= 1e-9 * max(abs(want_sum), 1.0)
    del keep
    for k in phases:
        phases[k] = 0.0
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        mat = e2e_step()
    torch.cuda.synchronize()
    e2e_bam_s = (time.perf_counter() - t0) / e2e_steps
    bam_phases = dict(phases)
    phases = main_phases

    # ---- N > 1: the exchanged matrix on rank 0 must hold every rank's rows ----
    exchange_ok = None
    if world > 1:
        barrier()
        saved = (out_box["ptr"], out_box["ld"])
        out_box["ptr"], out_box["ld"] = out_box["m"].data_ptr(), R      # one more step, local output
        device_step()
        out_box["ptr"], out_box["ld"] = saved
        if "peer" not in out_box:           # and one more gather of exactly that block
            gather_step(0)
        barrier()
        mine = out_box["m"].double().sum().reshape(1)
        sums = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(sums, mine)
        if rank == 0:
            full = out_box["peer"].as_tensor() if "peer" in out_box else gather_box["full"]
            exchange_ok = True
            for r in range(world):
                got = float(full[:, r * R:(r + 1) * R].sum())
                want = float(sums[r])
                if not (abs(got - want) <= 1e-9 * max(abs(want), 1.0)) or want == 0.0:
                    exchange_ok = False
            assert exchange_ok, "rows of some rank did not arrive in rank 0's matrix"

    # ---- reduce over ranks (max time) ----
    t = torch.tensor([elapsed_ms, e2e_s * 1e3, e2e_bam_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms, e2e_bam_ms = float(t[0]), float(t[1]), float(t[2])
    ms_per_step = elapsed_ms / args.steps
    value = world * N / (ms_per_step * 1e-3)
    e2e_value = world * N / (e2e_ms * 1e-3)

    if rank == 0:
        peak, peak_src = peaks()
        total_len = stats["total_len"]
        # Algorithmic bytes (SURVEY 8d / DESIGN.md 4).  Per STAGE:
        #   reads_map  13 B/read in (chrom, start, end, strand) + 9 B/read out (global start,
        #              end+1, strand)
        #   coverage   8 B/read + 4 B/covered base + 16 B/region     (everything between the
        #              mapped reads and the dense coverage: sort or bucketing included in the
        #              TIME, not in the bytes)
        #   profile    4 B/covered base + 8 B/matrix cell
        # Per KERNEL (the `roofline` object, dominant kernel of the step): the bytes that kernel
        # cannot avoid moving -- a pass over the reads 8 B/read (+1 with strand), the tile
        # kernels 4 B/covered base written (+ 8 B/read on the index path, which reads the sorted
        # arrays there), the bin kernel 4 B/covered base + 8 B/cell, the sort 8 B/key.  Block path:
        # the filter reads 8 B/read (its output, 8 B/candidate, is not counted); a scatter pass
        # reads and writes 8 B/candidate; the tile kernel writes 4 B/covered base and reads the
        # candidates once (8 B each; the re-reads of a block by its second tile are not counted).
        cand = stats.get("candidates", 0)
        alg_kernel = {
            "blk_filter": 8 * N,
            "blk_scatter": 16 * cand,
            "blk_tile": 4 * total_len + 8 * cand,
            "blk_small": 4 * total_len + 8 * cand,
            "sp_split": 8 * N,
            "sp_tile": 4 * total_len + 4 * cand,
            "sp_small": 4 * total_len + 4 * cand,
            "index_map": 22 * N,
            "index_sort": 8 * N,
            "cov_tile": 8 * N + 4 * total_len + 16 * R,
            "cov_small": 8 * N + 4 * total_len + 16 * R,
            "bkt_count": 8 * N,
            "bkt_scatter": 8 * N,
            "bkt_tile": 4 * total_len,
            "bkt_small": 4 * total_len,
            "prof_bin": 4 * total_len + 8 * R * ncols,
            "prof_base": 4 * total_len + 8 * R * ncols,
        }
        groups = {
            "reads_map": (["index_map"], 22 * N),
            "coverage": (["index_sort", "cov_plan", "cov_tile", "cov_small", "cov_list", "cov_concat",
                          "bkt_plan", "bkt_count", "bkt_scatter", "bkt_tile", "bkt_small",
                          "blk_filter", "blk_hist", "blk_scatter", "blk_tile", "blk_small",
                          "sp_plan", "sp_split", "sp_sort", "sp_tile", "sp_small"],
                         8 * N + 4 * total_len + 16 * R),
            "profile": (["prof_bin", "prof_interp", "prof_base"], 4 * total_len + 8 * R * ncols),
        }
        per_step = {k: v[0] * v[1] / args.steps for k, v in stage.items()}
        stage_roof = {}
        for g, (members, nbytes) in groups.items():
            ms_g = sum(per_step.get(m, 0.0) for m in members)
            if ms_g > 0:
                ach = nbytes / (ms_g * 1e-3) / 1e9
                stage_roof[g] = {"ms_per_step": ms_g, "algorithmic_bytes": nbytes, "achieved": ach,
                                 "frac": ach / peak}
        # measured DRAM traffic per launch (ncu --set full), available for the full-size C2 step
        traffic = {}
        tpath = os.path.join(ROOT, "profiles", "r01_traffic_C2.json")
        if args.workload == "C2" and args.scale == 1.0 and os.path.exists(tpath):
            traffic = json.load(open(tpath))
        own = {k: v for k, v in stage.items() if k in alg_kernel}
        dom = max(own, key=lambda k: own[k][0] * own[k][1]) if own else None
        roof = None
        if dom:
            per_step_ms = per_step[dom]
            per_launch_ms = own[dom][0]         # average over the launches of the timed region
            ach = alg_kernel[dom] / (per_launch_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic.get(dom), "peak_source": peak_src,
                    "ms_per_step": per_step_ms, "ms_per_launch": per_launch_ms,
                    "launches_per_step": own[dom][1] / args.steps,
                    "algorithmic_bytes": alg_kernel[dom],
                    "kernels": {k: {"ms_per_launch": own[k][0], "launches_per_step": own[k][1] / args.steps,
                                    "algorithmic_bytes": alg_kernel[k],
                                    "frac": alg_kernel[k] / (own[k][0] * 1e-3) / 1e9 / peak}
                                for k in own},
                    "stages": stage_roof}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32 coverage / f64 matrix",
            "data": "synthetic",
            "config": {"workload": w["name"], "regions_per_gpu": R, "reads_per_gpu": N,
                       "matrix_cols": ncols, "covered_bases_per_gpu": total_len,
                       "null_regions": stats["n_null"],
                       "l2": "inputs (%.0f MB) and coverage (%.0f MB) exceed the 126 MB L2"
                             % (13 * N / 1e6, 4 * total_len / 1e6),
                       "coverage_path": args.path, "coverage_path_used": stats.get("path"),
                       "filter_candidates_per_gpu": stats.get("candidates"),
                       "parallelism": ("regions sharded over %d GPU(s), " % world) +
                                      ("rows stored into rank 0's matrix over NVLink (peer-mapped)"
                                       if world > 1 and args.exchange == "p2p" else
                                       "NCCL row gather on its own stream (double-buffered matrix: "
                                       "the gather of step k runs beside step k + 1)")},
            "stage_ms_per_step": {k: v[0] * v[1] / args.steps for k, v in stage.items()},
            "region_bins_per_s": world * R * ncols / (ms_per_step * 1e-3),
            "roofline": roof,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "phases_ms": {k: 1e3 * v / e2e_steps for k, v in phases.items()},
                    "inputs": "reads in random order; seqnames, start, end int32 + strand int8"},
            "e2e_bam_order": {"value": world * N / (e2e_bam_ms * 1e-3), "unit": UNIT,
                              "h2d_bytes_per_step": int(h2d_bam), "d2h_bytes_per_step": int(d2h),
                              "ms_per_step": e2e_bam_ms, "steps": e2e_steps,
                              "phases_ms": {k: 1e3 * v / e2e_steps for k, v in bam_phases.items()},
                              "inputs": "the same reads grouped by chromosome (BAM order); seqnames "
                                        "as the Rle a GRanges holds (%d runs), start, end int32 + "
                                        "strand int8" % seq_rle.nrun},
            "gpu_launches": launches, "clocks": clocks,
        }
        if exchange_ok is not None:
            out["exchange_verified"] = exchange_ok
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(w)
        print(json.dumps(out))
    if world > 1:
        if gthread is not None:
            gq.put(None)
            gthread.join()
        if "peer" in out_box:
            out_box["peer"].close()
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
