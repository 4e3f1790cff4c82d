"""ctypes loader for oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never by rufus_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Table(C.Structure):
    _fields_ = [("keys", C.POINTER(C.c_uint64)), ("counts", C.POINTER(C.c_uint32)), ("n", C.c_uint64)]


def build():
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "rufus_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "oracle"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.ro_count.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.POINTER(_Table)]
        L.ro_count.restype = C.c_int
        L.ro_table_free.argtypes = [C.POINTER(_Table)]
        L.ro_histo_buckets.argtypes = [C.c_uint64, C.c_uint64]
        L.ro_histo_buckets.restype = C.c_uint64
        L.ro_histo.argtypes = [C.POINTER(_Table), C.c_uint64, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64)]
        L.ro_merge_unique.argtypes = [C.POINTER(C.POINTER(_Table)), C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]
        L.ro_merge_unique.restype = C.c_uint64
        L.ro_hashlist.argtypes = [C.POINTER(_Table), C.POINTER(C.POINTER(_Table)), C.c_int, C.c_uint64,
                                  C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        L.ro_hashlist.restype = C.c_uint64
        L.ro_rufus_pack.argtypes = [C.c_char_p, C.c_int]
        L.ro_rufus_pack.restype = C.c_uint64
        L.ro_rufus_revcomp.argtypes = [C.c_char_p, C.c_int, C.c_char_p]
        L.ro_rufus_revcomp.restype = C.c_int
        L.ro_filter_pairs.argtypes = [C.c_void_p] * 6 + [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                                         C.c_uint64, C.c_void_p]
        L.ro_filter_single.argtypes = [C.c_void_p] * 3 + [C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_uint64,
                                                          C.c_void_p]
        L.ro_slice_hash.argtypes = [C.c_uint64]
        L.ro_slice_hash.restype = C.c_uint32
        L.ro_slice_new.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint32]
        L.ro_slice_new.restype = C.c_void_p
        L.ro_slice_add.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32]
        L.ro_slice_finish.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(_Table)]
        L.ro_slice_finish.restype = C.c_int
        L.ro_slice_free.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


class SliceCounter:
    """k-mers of the slice {key : slice_hash(key) % mod == rem}, accumulated over chunks of fixed-length reads"""

    def __init__(self, k: int, canonical: bool, mod: int, rem: int):
        self.h = lib().ro_slice_new(k, int(canonical), mod, rem)
        assert self.h

    def add(self, reads: np.ndarray):
        r = np.ascontiguousarray(reads, np.uint8)
        lib().ro_slice_add(self.h, _p(r), r.shape[0], r.shape[1])

    def finish(self, lower: int = 1):
        t = _Table()
        assert lib().ro_slice_finish(self.h, lower, C.byref(t)) == 0
        keys = np.ctypeslib.as_array(t.keys, (t.n,)).copy() if t.n else np.zeros(0, np.uint64)
        counts = np.ctypeslib.as_array(t.counts, (t.n,)).copy() if t.n else np.zeros(0, np.uint32)
        lib().ro_table_free(C.byref(t))
        lib().ro_slice_free(self.h)
        self.h = None
        return keys, counts


def slice_hash(keys: np.ndarray) -> np.ndarray:
    """ro_slice_hash vectorised (the CUDA path's rg_hash32b: a fixed function of the key)"""
    k = np.asarray(keys, np.uint64)
    m32 = np.uint64(0xFFFFFFFF)
    h = ((k & m32) * np.uint64(0xC2B2AE35) + (k >> np.uint64(32)) * np.uint64(0x27D4EB2F)) & m32
    h ^= h >> np.uint64(15)
    return ((h * np.uint64(0x2C1B3C6D)) & m32).astype(np.uint32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _as_table(keys, counts):
    t = _Table()
    t.keys = keys.ctypes.data_as(C.POINTER(C.c_uint64))
    t.counts = counts.ctypes.data_as(C.POINTER(C.c_uint32))
    t.n = keys.size
    return t


def count(seq: np.ndarray, offs: np.ndarray, k: int, canonical: bool = True, lower: int = 1):
    """-> (keys u64 ascending, counts u32)."""
    seq = np.ascontiguousarray(seq, dtype=np.uint8).reshape(-1)
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    t = _Table()
    rc = lib().ro_count(_p(seq), _p(offs), offs.size - 1, k, int(canonical), lower, C.byref(t))
    if rc:
        raise RuntimeError(f"ro_count failed {rc}")
    keys = np.ctypeslib.as_array(t.keys, shape=(max(t.n, 1),))[: t.n].copy()
    counts = np.ctypeslib.as_array(t.counts, shape=(max(t.n, 1),))[: t.n].copy()
    lib().ro_table_free(C.byref(t))
    return keys, counts


def histo(counts: np.ndarray, low: int = 1, high: int = 10000):
    keys = np.zeros(counts.size, dtype=np.uint64)
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    t = _as_table(keys, counts)
    nb = lib().ro_histo_buckets(low, high)
    out = np.zeros(nb, dtype=np.uint64)
    base = C.c_uint64()
    lib().ro_histo(C.byref(t), low, high, _p(out), C.byref(base))
    return out, base.value


def merge_unique(tables, floor: int = 5):
    ts = [_as_table(np.ascontiguousarray(k, np.uint64), np.ascontiguousarray(c, np.uint32)) for k, c in tables]
    arr = (C.POINTER(_Table) * len(ts))(*[C.pointer(t) for t in ts])
    tot = sum(int(t.n) for t in ts)
    ok, oc = np.zeros(max(tot, 1), np.uint64), np.zeros(max(tot, 1), np.uint32)
    n = lib().ro_merge_unique(arr, len(ts), floor, _p(ok), _p(oc))
    return ok[:n].copy(), oc[:n].copy()


def hashlist(subj, others, min_cov: int, max_cov: int, floor: int = 5):
    ts = [_as_table(np.ascontiguousarray(k, np.uint64), np.ascontiguousarray(c, np.uint32)) for k, c in [subj] + list(others)]
    arr = (C.POINTER(_Table) * max(len(ts) - 1, 1))(*[C.pointer(t) for t in ts[1:]])
    tot = sum(int(t.n) for t in ts)
    ok, oc = np.zeros(max(tot, 1), np.uint64), np.zeros(max(tot, 1), np.uint32)
    n = lib().ro_hashlist(C.byref(ts[0]), arr, len(ts) - 1, floor, min_cov, max_cov, _p(ok), _p(oc))
    return ok[:n].copy(), oc[:n].copy()


def rufus_set_from_kmers(kmers) -> np.ndarray:
    """RUFUS.Filter.cpp:121-143 for 2-column lines: packed k-mer and packed RevComp, sorted unique."""
    L = lib()
    vals = []
    for s in kmers:
        b = s.encode() if isinstance(s, str) else bytes(s)
        vals.append(L.ro_rufus_pack(b, len(b)))
        out = C.create_string_buffer(len(b) + 1)
        n = L.ro_rufus_revcomp(b, len(b), out)
        vals.append(L.ro_rufus_pack(out.raw[:n], n))
    return np.unique(np.array(vals, dtype=np.uint64))


def filter_pairs(s1, q1, o1, s2, q2, o2, k, minq, thresh, rset):
    a = [np.ascontiguousarray(x, np.uint8).reshape(-1) for x in (s1, q1)]
    b = [np.ascontiguousarray(x, np.uint8).reshape(-1) for x in (s2, q2)]
    o1 = np.ascontiguousarray(o1, np.uint64); o2 = np.ascontiguousarray(o2, np.uint64)
    rset = np.ascontiguousarray(rset, np.uint64)
    keep = np.zeros(o1.size - 1, np.uint8)
    lib().ro_filter_pairs(_p(a[0]), _p(a[1]), _p(o1), _p(b[0]), _p(b[1]), _p(o2), o1.size - 1, k, minq, thresh,
                          _p(rset), rset.size, _p(keep))
    return keep


def filter_single(s, q, o, k, minq, rset):
    s = np.ascontiguousarray(s, np.uint8).reshape(-1); q = np.ascontiguousarray(q, np.uint8).reshape(-1)
    o = np.ascontiguousarray(o, np.uint64); rset = np.ascontiguousarray(rset, np.uint64)
    hits = np.zeros(o.size - 1, np.int32)
    lib().ro_filter_single(_p(s), _p(q), _p(o), o.size - 1, k, minq, _p(rset), rset.size, _p(hits))
    return hits


def build_select(subj, controls, min_subject: int, max_subject: int, max_control_sum: int):
    """RUFUS.Build (src/RUFUS.Build.cpp:224-262) restated on sorted (keys, counts) tables: the subject
    entries with min_subject <= count <= max_subject (:226) whose summed count over the control tables
    (:245-252, a control contributes only where the k-mer is present) is <= max_control_sum (:257).
    Returns (keys, subject counts, control sums) in key order = the order of the sorted text dump."""
    sk, sc = subj
    sel = (sc >= min_subject) & (sc <= max_subject)
    keys, counts = sk[sel], sc[sel]
    total = np.zeros(keys.size, np.int64)
    for ck, cc in controls:
        if ck.size == 0:
            continue                      # an empty control contributes nothing
        pos = np.searchsorted(ck, keys)
        pos[pos >= ck.size] = 0
        hit = ck[pos] == keys
        total += np.where(hit, cc[pos].astype(np.int64), 0)
    keep = total <= max_control_sum
    return keys[keep], counts[keep], total[keep].astype(np.uint32)


def annotate(seq, qual, offs, k: int, rset, minq: int = 3) -> np.ndarray:
    """AnnotateOverlap (src/AnnotateOverlap.cpp:95-146) restated: per-base depth of mutant k-mer windows.
    Window i of a read of length L is examined for i < L - k (:101), counts when it holds no 'N' and every
    quality-33 >= minq (:104-115) and its rufus key or its reverse complement's is in the set (:117-128;
    `rset` holds k-mer and reverse complement, like the tool's two lookups). Small inputs only."""
    seq = np.ascontiguousarray(seq, np.uint8).reshape(-1)
    qual = np.ascontiguousarray(qual, np.uint8).reshape(-1)
    offs = np.ascontiguousarray(offs, np.uint64)
    members = set(int(x) for x in np.asarray(rset, np.uint64).tolist())
    code = {ord("C"): 2, ord("G"): 1, ord("T"): 3}
    depth = np.zeros(seq.size, np.uint8)
    for r in range(offs.size - 1):
        a, b = int(offs[r]), int(offs[r + 1])
        s, q = seq[a:b], qual[a:b]
        for i in range(0, max(0, (b - a) - k)):
            w, wq = s[i:i + k], q[i:i + k]
            if (w == ord("N")).any() or (wq.astype(np.int64).astype(np.int8).astype(np.int64) - 33 < minq).any():
                continue
            v = 0
            for j, c in enumerate(w.tolist()):
                v |= code.get(c, 0) << (2 * j)
            if v in members:
                d = depth[a + i:a + i + k].astype(np.int64) + 1
                depth[a + i:a + i + k] = np.minimum(d, 255).astype(np.uint8)
    return depth
