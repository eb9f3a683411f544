"""Test helper: writes BAM files (SAMv1 section 4) from python lists, so that the decoder tests do
not depend on samtools.  Independent of the parsers it is used against (oracle/import_oracle.py,
the CUDA decoder): it only ever WRITES the format."""
import struct
import zlib

OPS = "MIDNSHP=X"


def bgzf_block(data):
    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = comp.compress(data) + comp.flush()
    bsize = len(body) + 25                      # total block size - 1
    head = struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize)
    return head + body + struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data))


def bgzf_compress(raw, block=60000):
    out = [bgzf_block(raw[i:i + block]) for i in range(0, len(raw), block)]
    out.append(bgzf_block(b""))                 # the EOF marker
    return b"".join(out)


def cigar_ops(cigar):
    """"10M2I5N3M" -> [(10, 'M'), ...]"""
    ops, num = [], ""
    for ch in cigar:
        if ch.isdigit():
            num += ch
        else:
            ops.append((int(num), ch))
            num = ""
    return ops


def bam_record(ref, pos0, flag, cigar, name=b"r", l_seq=0, tags=b""):
    """One alignment record; pos0 is 0-based (the file's own convention)."""
    ops = cigar_ops(cigar) if cigar != "*" else []
    name = name + b"\x00"
    body = struct.pack("<iiBBHHHiiii", ref, pos0, len(name), 30, 4680, len(ops), flag, l_seq, -1, -1, 0)
    body += name
    body += b"".join(struct.pack("<I", (ln << 4) | OPS.index(op)) for ln, op in ops)
    body += b"\x00" * ((l_seq + 1) // 2) + b"\xff" * l_seq + tags
    return struct.pack("<i", len(body)) + body


def bam_file(refs, records, text=b"@HD\tVN:1.0\tSO:unsorted\n"):
    """refs: [(name, length)]; records: bytes objects from bam_record.  Returns (inflated, bgzf)."""
    raw = b"BAM\x01" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs))
    for name, ln in refs:
        nm = name.encode("ascii") + b"\x00"
        raw += struct.pack("<i", len(nm)) + nm + struct.pack("<i", ln)
    raw += b"".join(records)
    return raw, bgzf_compress(raw)
