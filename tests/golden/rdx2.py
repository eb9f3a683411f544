"""Minimal reader for R's `save()` format (gzip + "RDX2\\nX\\n", XDR serialisation v2).

Test infrastructure only: used by `make_fixtures.py` to decode the reference's bundled
`data/recoup_test_data.rda` (documented in /root/reference/man/recoup_test_data.Rd:1-39) into
plain integer arrays.  Nothing in the product imports this.

Only the SEXP types that occur in that file are handled.  S4 objects come back as
`{"__class__": ..., "<slot>": ...}` dictionaries (slots are stored as an attribute pairlist).
"""
import gzip
import struct

import numpy as np

NILVALUE, REFSXP, GLOBALENV, EMPTYENV, BASEENV = 254, 255, 253, 242, 241
NAMESPACESXP, PACKAGESXP, PERSISTSXP, MISSINGARG, UNBOUND = 249, 250, 247, 251, 252
ATTRLANGSXP, ATTRLISTSXP, BASENAMESPACE = 240, 239, 247


class _Reader:
    def __init__(self, buf):
        self.b = buf
        self.p = 0
        self.refs = []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def raw(self, n):
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if t == NILVALUE:
            return None
        if t in (GLOBALENV, EMPTYENV, BASEENV, MISSINGARG, UNBOUND):
            return None
        if t == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == 1:  # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t in (NAMESPACESXP, PACKAGESXP, PERSISTSXP):
            self.i32()
            n = self.i32()
            v = [self.item() for _ in range(n)]
            self.refs.append(v)
            return v
        if t in (2, 6, ATTRLISTSXP, ATTRLANGSXP):  # pairlist / language
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.i32()
                t2 = flags & 0xFF
                if t2 == NILVALUE:
                    break
                if t2 not in (2, 6, ATTRLISTSXP, ATTRLANGSXP):
                    raise ValueError("improper pairlist tail type %d" % t2)
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
            return out
        if t == 9:  # CHARSXP
            n = self.i32()
            return None if n == -1 else self.raw(n).decode("latin1")
        if t in (10, 13):  # LGLSXP / INTSXP
            n = self.i32()
            v = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.p).astype(np.int32)
            self.p += 4 * n
        elif t == 14:  # REALSXP
            n = self.i32()
            v = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.p).astype(np.float64)
            self.p += 8 * n
        elif t == 16:  # STRSXP
            n = self.i32()
            v = [self.item() for _ in range(n)]
        elif t in (19, 20):  # VECSXP / EXPRSXP
            n = self.i32()
            v = [self.item() for _ in range(n)]
        elif t == 25:  # S4SXP
            v = {}
        elif t == 24:  # RAWSXP
            n = self.i32()
            v = self.raw(n)
        else:
            raise ValueError("unsupported SEXP type %d at %d" % (t, self.p))
        if has_attr:
            attrs = dict((k, val) for k, val in self.item())
            if isinstance(v, dict):
                cls = attrs.pop("class", None)
                v.update(attrs)
                v["__class__"] = val(cls)[0] if cls is not None else None
            else:
                v = _Attributed(v, attrs)
        return v


class _Attributed:
    """A vector plus its R attributes (names, levels, class, row.names ...)."""

    def __init__(self, value, attrs):
        self.value = value
        self.attrs = attrs

    def __repr__(self):
        return "Attributed(%r, attrs=%s)" % (type(self.value).__name__, list(self.attrs))


def val(x):
    return x.value if isinstance(x, _Attributed) else x


def attr(x, name):
    return x.attrs.get(name) if isinstance(x, _Attributed) else None


def load_rda(path):
    buf = gzip.open(path, "rb").read()
    if buf[:5] != b"RDX2\n" or buf[5:7] != b"X\n":
        raise ValueError("not an RDX2/XDR file")
    r = _Reader(buf)
    r.p = 7
    version, _writer, _minr = r.i32(), r.i32(), r.i32()
    if version != 2:
        raise ValueError("serialisation version %d not supported" % version)
    top = r.item()  # a pairlist of (name, object)
    return dict(top)
