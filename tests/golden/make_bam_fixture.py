#!/usr/bin/env python
"""Cut a small real-world BAM out of the reference's bundled sample.

Run in the build container (the only place /root/reference exists):

    python tests/golden/make_bam_fixture.py

Reads /root/reference/inst/extdata/WT_H4K20me1_50kr.bam (the file the reference's vignette feeds to
readBam, ranges.R:111-134) and writes `tests/golden/WT_H4K20me1_5kr.bam`: the same header and the
first 5 000 alignment records, re-compressed as BGZF.  Data only -- it gives the BAM decoder a file
written by a real tool (bedToBam) next to the synthetic ones the tests build themselves.  The
reference holds no decoded output for it, so it pins INPUT handling, not values (the values are
checked against oracle/import_oracle.py)."""
import os
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import import_oracle as IO  # noqa: E402
from tests.bam_writer import bgzf_compress  # noqa: E402

SRC = "/root/reference/inst/extdata/WT_H4K20me1_50kr.bam"
N = 5000

raw = IO.bgzf_inflate(open(SRC, "rb").read())
_, _, first = IO.bam_header(raw)
p = first
for _ in range(N):
    bs, = struct.unpack_from("<i", raw, p)
    p += 4 + bs
out = os.path.join(HERE, "WT_H4K20me1_5kr.bam")
open(out, "wb").write(bgzf_compress(raw[:p]))
print(out, os.path.getsize(out), "bytes")
