"""Input side of ``run_mcmc``: the reference's data loaders and a reader for its packaged ``.rda``
files, restated in Python (R is not available where the engine runs).

* ``load_long_form_data`` / ``load_delimited_data`` follow R/utils.R:17-84 and :102-111: drop loci
  with a single allele, alleles sorted per locus, samples sorted, one 0/1 barcode per
  (locus, sample), ``is_missing`` as a loci x samples matrix.  The result has the fields the R
  functions return and feeds ``moire_b200.mcmc.run_mcmc`` / ``moire_b200.data.pack_data``.
* ``read_rda`` decodes R's serialisation (``RDX2``/``RDX3`` header, XDR, xz / bz2 / gzip
  compressed): the SEXP types the packaged data files use.

Host-side preparation only; nothing here is on the likelihood path.
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct
from collections import OrderedDict

import numpy as np


# ---- loaders (R/utils.R) ---------------------------------------------------------------------------
def _columns(df, names):
    if hasattr(df, "columns") and hasattr(df, "__getitem__") and not isinstance(df, dict):
        return [list(df[n]) for n in names]          # pandas-like
    return [list(df[n]) for n in names]


def _r_sort_key(x):
    # R's sort() on character vectors in the C locale tests of the reference: plain string order;
    # numbers sort numerically
    return (0, x) if isinstance(x, (int, float, np.integer, np.floating)) else (1, str(x))


def load_long_form_data(df, warn_uninformative=True):
    """``df``: columns ``sample_id``, ``locus``, ``allele`` (pandas DataFrame or dict of
    sequences), one row per observed allele.  Returns ``dict(sample_ids, data, loci, is_missing,
    uninformative_loci)`` with ``data[locus]`` an int8 [N][A_locus] 0/1 matrix (the reference's
    ``data[[locus]][[sample]]`` barcodes stacked) and ``is_missing`` bool [L][N]."""
    sid, loc, al = _columns(df, ["sample_id", "locus", "allele"])
    if not (len(sid) == len(loc) == len(al)):
        raise ValueError("sample_id, locus and allele must have the same length")
    alleles_of = {}
    for l, a in zip(loc, al):
        alleles_of.setdefault(l, set()).add(a)
    uninformative = sorted((l for l, s in alleles_of.items() if len(s) == 1), key=_r_sort_key)
    if uninformative:
        if warn_uninformative:
            print("Uninformative loci with only 1 allele included. Removing...")
        keep = [i for i, l in enumerate(loc) if l not in set(uninformative)]
        sid, loc, al = [sid[i] for i in keep], [loc[i] for i in keep], [al[i] for i in keep]
        for l in uninformative:
            del alleles_of[l]
    loci = sorted(alleles_of, key=_r_sort_key)
    sample_ids = sorted(set(sid), key=_r_sort_key)
    s_index = {s: i for i, s in enumerate(sample_ids)}
    N, L = len(sample_ids), len(loci)
    data, is_missing = [], np.ones((L, N), dtype=bool)
    a_index = []
    for j, l in enumerate(loci):
        order = sorted(alleles_of[l], key=_r_sort_key)
        a_index.append({a: k for k, a in enumerate(order)})
        data.append(np.zeros((N, len(order)), dtype=np.int8))
    l_index = {l: j for j, l in enumerate(loci)}
    for s, l, a in zip(sid, loc, al):
        j, i = l_index[l], s_index[s]
        data[j][i, a_index[j][a]] = 1
        is_missing[j, i] = False
    return dict(sample_ids=sample_ids, data=data, loci=loci, is_missing=is_missing,
                uninformative_loci=uninformative)


def load_delimited_data(data, sep=";", warn_uninformative=True):
    """``data``: a ``sample_id`` column plus one column per locus whose cells hold the observed
    alleles joined by ``sep`` (None / NaN = nothing observed).  R/utils.R:102-111."""
    cols = list(data.columns) if hasattr(data, "columns") else list(data.keys())
    sid_col = list(data["sample_id"])
    sid, loc, al = [], [], []
    for c in cols:
        if c == "sample_id":
            continue
        for s, cell in zip(sid_col, list(data[c])):
            if cell is None or (isinstance(cell, float) and np.isnan(cell)):
                continue
            for a in str(cell).split(sep):
                sid.append(s)
                loc.append(c)
                al.append(a)
    return load_long_form_data(dict(sample_id=sid, locus=loc, allele=al), warn_uninformative)


# ---- R serialisation reader ------------------------------------------------------------------------
class RObject:
    """A decoded R value that carries attributes the caller may need (names, class, levels...)."""

    def __init__(self, value, attrs):
        self.value, self.attrs = value, attrs

    def __repr__(self):
        return f"RObject({type(self.value).__name__}, attrs={list(self.attrs)})"


class _XdrReader:
    NA_INT = -2 ** 31

    def __init__(self, buf):
        self.b, self.o, self.refs = buf, 0, []

    def int(self):
        v = struct.unpack_from(">i", self.b, self.o)[0]
        self.o += 4
        return v

    def length(self):
        n = self.int()
        if n == -1:                      # long vector: two more ints
            hi, lo = self.int(), self.int()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def bytes(self, n):
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def item(self):
        flags = self.int()
        t = flags & 0xFF
        has_attr, has_tag = bool(flags & (1 << 9)), bool(flags & (1 << 10))
        if t == 254:                                     # NILVALUE
            return None
        if t in (253, 242, 241, 251, 252, 248):          # environments / missing / unbound markers
            return None
        if t == 255:                                     # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.int()
            return self.refs[idx - 1]
        if t == 1:                                       # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t in (249, 250, 247):                         # NAMESPACE / PACKAGE / PERSIST
            self.int()
            info = [self.item() for _ in range(self.int())] if t != 247 else self.item()
            self.refs.append(info)
            return info
        if t == 4:                                       # ENVSXP: locked, enclos, frame, hashtab, attr
            self.int()
            self.refs.append(None)
            for _ in range(4):
                self.item()
            return None
        if t in (2, 6, 3, 5, 17, 239, 240):              # pairlist-like
            out = []
            while True:
                attrs = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.int()
                t2 = flags & 0xFF
                if t2 == 254:
                    break
                if t2 not in (2, 6, 3, 5, 17, 239, 240):
                    # the cdr is some other object: rewind and read it as an item
                    self.o -= 4
                    out.append((None, self.item()))
                    break
                has_attr, has_tag = bool(flags & (1 << 9)), bool(flags & (1 << 10))
            return out
        if t == 9:                                       # CHARSXP
            n = self.int()
            return None if n == -1 else self.bytes(n).decode("utf-8", "replace")
        if t == 238:                                     # ALTREP
            info, state, attr = self.item(), self.item(), self.item()
            cls = info[0][1] if info else None
            val = self._altrep(cls, state)
            return self._with_attrs(val, attr)
        if t in (10, 13):                                # LGLSXP / INTSXP
            n = self.length()
            v = np.frombuffer(self.bytes(4 * n), dtype=">i4").astype(np.int64)
            if t == 10:
                v = np.where(v == self.NA_INT, -1, v).astype(np.int8)   # NA -> -1
        elif t == 14:                                    # REALSXP
            n = self.length()
            v = np.frombuffer(self.bytes(8 * n), dtype=">f8").astype(np.float64)
        elif t == 15:                                    # CPLXSXP
            n = self.length()
            v = np.frombuffer(self.bytes(16 * n), dtype=">c16").astype(np.complex128)
        elif t == 16:                                    # STRSXP
            v = [self.item() for _ in range(self.length())]
        elif t in (19, 20):                              # VECSXP / EXPRSXP
            v = [self.item() for _ in range(self.length())]
        elif t == 24:                                    # RAWSXP
            v = bytes(self.bytes(self.length()))
        elif t in (8, 7):                                # BUILTIN / SPECIAL
            v = self.bytes(self.int()).decode()
        elif t == 25:                                    # S4SXP
            v = None
        else:
            raise ValueError(f"unsupported SEXP type {t} at offset {self.o}")
        attrs = self.item() if has_attr else None
        return self._with_attrs(v, attrs)

    def _altrep(self, cls, state):
        if cls in ("compact_intseq", "compact_realseq"):
            n, start, step = (float(x) for x in (state.value if isinstance(state, RObject) else state))
            seq = start + step * np.arange(int(n))
            return seq.astype(np.int64) if cls == "compact_intseq" else seq
        if cls == "deferred_string":
            arg = state[0][1]
            arg = arg.value if isinstance(arg, RObject) else arg
            return [str(int(x)) if float(x).is_integer() else str(x) for x in np.asarray(arg)]
        if cls and cls.startswith("wrap_"):
            inner = state[0] if isinstance(state, list) else state
            if isinstance(inner, tuple):
                inner = inner[1]
            return inner.value if isinstance(inner, RObject) else inner
        raise ValueError(f"unsupported ALTREP class {cls!r}")

    @staticmethod
    def _with_attrs(v, attrs):
        if not attrs:
            return v
        a = OrderedDict()
        for tag, val in attrs:
            a[tag] = val.value if isinstance(val, RObject) and not val.attrs else val
        if isinstance(v, np.ndarray) and "levels" in a and v.dtype.kind == "i":   # factor -> strings
            lv = a["levels"]
            return [None if x < 1 else lv[int(x) - 1] for x in v]
        if isinstance(v, list) and "names" in a and a["names"] is not None and \
                len(set(a["names"])) == len(v) and set(a) <= {"names", "class", "row.names"}:
            return OrderedDict(zip(a["names"], v))          # named list / data frame
        if isinstance(v, np.ndarray) and "dim" in a and set(a) <= {"dim", "dimnames"}:
            return v.reshape([int(d) for d in np.asarray(a["dim"])], order="F")
        if isinstance(v, np.ndarray) and set(a) <= {"names"}:
            return v
        return RObject(v, a)


def read_rda(path):
    """All objects saved in an ``.rda`` / ``.RData`` file: ``{name: value}``.  Vectors come back
    as numpy arrays (logical as int8 with NA = -1), character vectors as lists, factors as lists of
    strings, named lists and data frames as ordered dicts, matrices reshaped column-major."""
    raw = open(path, "rb").read()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    elif raw[:3] == b"BZh":
        raw = bz2.decompress(raw)
    elif raw[:6] == b"\xfd7zXZ\x00":
        raw = lzma.decompress(raw)
    if raw[:5] not in (b"RDX2\n", b"RDX3\n"):
        raise ValueError("not an R data file (RDX2/RDX3 header missing)")
    if raw[5:7] != b"X\n":
        raise ValueError("only the XDR (binary big-endian) serialisation is supported")
    r = _XdrReader(raw)
    r.o = 7
    version = r.int()
    r.int()
    r.int()
    if version == 3:
        r.bytes(r.int())                 # native encoding name
    top = r.item()
    return OrderedDict((tag, val) for tag, val in top)
