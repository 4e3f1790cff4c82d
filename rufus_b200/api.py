"""ctypes binding of librufus_gpu.so (include/rufus_gpu.h) + a thin host-side
mirror of the reference's stage commands for this path:

  KmerTable.count(...)      jellyfish count -m K -C            (RunJellyForRUFUS.sh:29)
  KmerTable.histo()         jellyfish histo -f                 (RunJellyForRUFUS.sh:37)
  KmerTable.dump()          jellyfish dump -c                  (Overlap.shorter.sh:247)
  KmerTable.query()         jellyfish query                    (CheckJellyHashList.sh:12)
  merge_unique(), subtract() modified `jellyfish merge` (+ query + awk cut-offs, runRufus.sh:925-926)
  MutantSet / filter_pairs / filter_single   RUFUS.Filter, RUFUS.Filter.single (runRufus.sh:967,981)

There is no CPU path: importing works anywhere (so the symbol check can run
without a GPU), but every compute call needs the CUDA library and a device and
raises RufusGpuError otherwise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librufus_gpu.so")

MODE_AUTO, MODE_DIRECT, MODE_TILED = 0, 1, 2
ERR_SMALL = -5


class RufusGpuError(RuntimeError):
    pass


_lib = None


def lib():
    """Load librufus_gpu.so; fail loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    global LIB_PATH
    if os.environ.get("RG_LIB"):           # development: an alternative build of the library (A/B kernel variants)
        LIB_PATH = os.path.abspath(os.environ["RG_LIB"])
    if not os.path.exists(LIB_PATH):
        raise RufusGpuError(f"{LIB_PATH} is missing: run `make lib` (or __graft_entry__.build()); "
                            "rufus_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i32, i64 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_int64
    sig = {
        "rg_last_error": (C.c_char_p, []),
        "rg_version": (C.c_char_p, []),
        "rg_init": (vp, [i32]),
        "rg_shutdown": (None, [vp]),
        "rg_set_stream": (i32, [vp, vp]),
        "rg_sync": (i32, [vp]),
        "rg_launch_count": (u64, [vp]),
        "rg_dev_alloc": (vp, [vp, u64]),
        "rg_dev_free": (None, [vp, vp]),
        "rg_host_alloc": (vp, [u64]),
        "rg_host_free": (None, [vp]),
        "rg_h2d": (i32, [vp, vp, vp, u64]),
        "rg_d2h": (i32, [vp, vp, vp, u64]),
        "rg_count_begin": (vp, [vp, i32, i32, u64, i32]),
        "rg_count_push": (i32, [vp, vp, u64]),
        "rg_count_push_reads": (i32, [vp, vp, vp, u64]),
        "rg_count_push_dev": (i32, [vp, vp, u64, u32]),
        "rg_count_end": (i32, [vp]),
        "rg_table_reserve": (i32, [vp, u64]),
        "rg_table_free": (None, [vp]),
        "rg_table_clear": (i32, [vp]),
        "rg_table_trim": (i32, [vp]),
        "rg_profile_enable": (i32, [vp, i32]),
        "rg_profile_reset": (i32, [vp]),
        "rg_profile_get": (i32, [vp, C.c_char_p, C.POINTER(C.c_double), C.POINTER(u64)]),
        "rg_profile_names": (i32, [vp, C.c_char_p, u64]),
        "rg_table_stats": (i32, [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]),
        "rg_table_k": (i32, [vp]),
        "rg_table_canonical": (i32, [vp]),
        "rg_extract_partition_dev": (i32, [vp, vp, u64, u32, vp, vp]),
        "rg_key_shard": (u32, [u64, i32, i32, u32, u32]),
        "rg_device_count": (i32, []),
        "rg_device_memory": (i32, [vp, vp, vp]),
        "rg_dump_slice_dev": (i64, [vp, u64, u32, u32, vp, vp, u64, vp]),
        "rg_synth_pairs_dev": (i32, [vp, vp, vp, vp, vp, vp]),
        "rg_synth_pairs_host": (i32, [vp, vp, vp, vp, vp]),
        "rg_tables_p2p_attach": (i32, [vp, i32]),
        "rg_tables_write_jhash": (i64, [vp, i32, C.c_char_p, u64, u64, u64, u32, vp, i32]),
        "rg_defer_checks": (i32, [vp, i32]),
        "rg_table_status_async": (i32, [vp, vp]),
        "rg_table_status_dev": (i32, [vp, vp]),
        "rg_histo_dev": (i32, [vp, u64, u64, u64, vp]),
        "rg_subtract_enqueue": (i32, [vp, u64, vp, vp, i32, u64, i64, i64, vp, vp, u64, vp]),
        "rg_set_create": (vp, [vp, i32, u64]),
        "rg_set_rebuild_dev": (i32, [vp, vp, u64]),
        "rg_minimizer_len": (i32, [i32, u64]),
        "rg_table_geometry": (i32, [vp, vp, vp, vp, vp]),
        "rg_count_push_keys_dev": (i32, [vp, vp, u64]),
        "rg_table_set_shard": (i32, [vp, u32, u32]),
        "rg_table_p2p_prepare": (i32, [vp, u32, u64]),
        "rg_table_p2p_export": (i32, [vp, vp]),
        "rg_table_p2p_import": (i32, [vp, u32, vp]),
        "rg_count_p2p_begin": (i32, [vp]),
        "rg_count_p2p_scatter_dev": (i32, [vp, vp, u64, u32, u32]),
        "rg_use_stream": (i32, [vp, vp]),
        "rg_count_p2p_finish": (i32, [vp]),
        "rg_histo_rows": (u64, [u64, u64]),
        "rg_histo": (i32, [vp, u64, u64, u64, vp]),
        "rg_dump": (i64, [vp, u64, u64, vp, vp, u64, C.POINTER(u64)]),
        "rg_table_from_pairs": (vp, [vp, i32, i32, vp, vp, u64]),
        "rg_query": (i32, [vp, vp, u64, u64, vp]),
        "rg_jf_matrix": (i32, [u64, i32, vp, C.POINTER(u32), C.POINTER(u64)]),
        "rg_jf_pos": (u64, [u64, u64, i32]),
        "rg_jf_sort": (i64, [vp, vp, vp, u64, u64, i32]),
        "rg_table_write_jhash": (i64, [vp, C.c_char_p, u64, u64, u64, u32, vp, i32]),
        "rg_table_read_jhash": (vp, [vp, C.c_char_p, vp]),
        "rg_merge_unique": (i64, [vp, vp, i32, u64, vp, vp, vp, u64, C.POINTER(u64)]),
        "rg_subtract": (i64, [vp, u64, vp, vp, i32, u64, i64, i64, vp, vp, u64, C.POINTER(u64)]),
        "rg_subtract_dev": (i64, [vp, u64, vp, vp, i32, u64, i64, i64, vp, vp, u64]),
        "rg_build_select": (i64, [vp, i64, i64, vp, i32, i64, vp, vp, vp, u64, C.POINTER(u64)]),
        "rg_rufus_pack": (u64, [C.c_char_p, i32]),
        "rg_rufus_revcomp": (i32, [C.c_char_p, i32, C.c_char_p]),
        "rg_set_from_hashlist_text": (vp, [vp, C.c_char_p, u64, i32, i32]),
        "rg_set_from_rufus_keys": (vp, [vp, vp, u64, i32]),
        "rg_set_from_jf_keys": (vp, [vp, vp, u64, i32]),
        "rg_set_from_jf_keys_dev": (vp, [vp, vp, u64, i32]),
        "rg_set_size": (u64, [vp]),
        "rg_set_free": (None, [vp]),
        "rg_filter_hits_dev": (i32, [vp, vp, vp, vp, u64, u64, i32, i32, vp]),
        "rg_filter_hits_fixed_dev": (i32, [vp, vp, vp, u32, u64, i32, i32, vp]),
        "rg_filter_pairs_fixed_dev": (i32, [vp, vp, vp, vp, vp, u32, u64, i32, i32, vp, vp, vp]),
        "rg_filter_pairs": (i32, [vp, vp, vp, vp, vp, vp, vp, u64, i32, i32, vp]),
        "rg_filter_single": (i32, [vp, vp, vp, vp, u64, i32, vp]),
        "rg_overlap_create": (vp, [vp, i32]),
        "rg_overlap_free": (None, [vp]),
        "rg_overlap_build": (i32, [vp, vp, vp, u64, u64]),
        "rg_overlap_update": (i32, [vp, u64, vp, u64]),
        "rg_overlap_seed_hits": (i64, [vp, vp, vp, vp, u64, vp, u64, vp]),
        "rg_overlap_align": (i32, [vp, vp, u64, vp, u64, C.c_float, i32, vp]),
        "rg_overlap_align_both": (i32, [vp, vp, vp, u64, vp, u64, vp, u64, C.c_float, i32, vp, vp]),
        "rg_annotate": (i32, [vp, vp, vp, vp, u64, i32, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


class JfInfo(C.Structure):
    """rg_jf_info: header values of a .Jhash file"""
    _fields_ = [("k", C.c_int), ("canonical", C.c_int), ("size", C.c_uint64), ("counter_len", C.c_uint32),
                ("n_records", C.c_uint64), ("matrix_r", C.c_uint32), ("max_reprobe_offset", C.c_uint64),
                ("columns", C.c_uint64 * 64)]


def jf_matrix(size: int, k: int):
    """(columns[2k], r, size rounded to a power of two) of a `jellyfish count -s size -m k` table"""
    cols = (C.c_uint64 * 64)()
    r, sz = C.c_uint32(), C.c_uint64()
    _check(lib().rg_jf_matrix(size, k, cols, C.byref(r), C.byref(sz)), "rg_jf_matrix")
    return [int(x) for x in cols[: 2 * k]], int(r.value), int(sz.value)


def _err(what: str):
    return RufusGpuError(f"{what}: {lib().rg_last_error().decode()}")


def _check(rc: int, what: str):
    if rc != 0:
        raise _err(what)


def _ptr(a):
    """numpy array -> host pointer; int -> raw (device) pointer; torch tensor -> data_ptr."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8).reshape(-1)


class Context:
    """One per GPU / process (rg_init)."""

    def __init__(self, device: int = 0):
        self.h = lib().rg_init(device)
        if not self.h:
            raise _err("rg_init")
        self.device = device

    def use_stream(self, cuda_stream_ptr: int | None):
        """Run on the caller's CUDA stream. None = the library's own stream; 0 = the legacy default
        stream (what torch.cuda.current_stream().cuda_stream reports for torch's default stream),
        passed to the library as cudaStreamLegacy (0x1) since a NULL handle means "library stream"."""
        if cuda_stream_ptr is None:
            p = 0
        else:
            p = cuda_stream_ptr or 1
        _check(lib().rg_set_stream(self.h, C.c_void_p(p)), "rg_set_stream")

    def switch_stream(self, cuda_stream_ptr: int):
        """use_stream without waiting for the previous stream's work (the caller orders them with events)"""
        _check(lib().rg_use_stream(self.h, C.c_void_p(cuda_stream_ptr or 1)), "rg_use_stream")

    def sync(self):
        _check(lib().rg_sync(self.h), "rg_sync")

    def profile(self, on: bool):
        _check(lib().rg_profile_enable(self.h, int(on)), "rg_profile_enable")

    def profile_reset(self):
        _check(lib().rg_profile_reset(self.h), "rg_profile_reset")

    def profile_report(self) -> dict:
        """kernel name -> (total ms, launches) since the last reset"""
        buf = C.create_string_buffer(4096)
        _check(lib().rg_profile_names(self.h, buf, 4096), "rg_profile_names")
        out = {}
        for name in buf.value.decode().split("\n"):
            if not name:
                continue
            ms, n = C.c_double(), C.c_uint64()
            if lib().rg_profile_get(self.h, name.encode(), C.byref(ms), C.byref(n)) == 0:
                out[name] = (ms.value, n.value)
        return out

    @property
    def launches(self) -> int:
        return int(lib().rg_launch_count(self.h))

    def close(self):
        if self.h:
            lib().rg_shutdown(self.h)
            self.h = None


class SynthSpec(C.Structure):
    """rg_synth_spec (include/rufus_gpu.h): the counter-based synthetic read generator's coordinates"""
    _fields_ = [("seed", C.c_uint64), ("genome_len", C.c_uint64), ("first_pair", C.c_uint64), ("n_pairs", C.c_uint64),
                ("read_len", C.c_uint32), ("sample", C.c_uint32), ("kind", C.c_uint32), ("repeats", C.c_uint32)]


SYNTH_TRIO, SYNTH_PAIR = 0, 1


def synth_pairs_host(spec: SynthSpec):
    """(seq1, qual1, seq2, qual2) uint8 [n_pairs, read_len], generated on the CPU (same bytes as the device)"""
    shape = (int(spec.n_pairs), int(spec.read_len))
    out = [np.empty(shape, np.uint8) for _ in range(4)]
    _check(lib().rg_synth_pairs_host(C.byref(spec), *[_ptr(a) for a in out]), "rg_synth_pairs_host")
    return tuple(out)


class KmerTable:
    """A counted k-mer table resident in HBM (the in-memory form of a .Jhash)."""

    def __init__(self, ctx: Context, k: int, canonical: bool = True, est_distinct: int = 0, mode: int = MODE_AUTO,
                 shard: tuple[int, int] | None = None, _handle=None):
        self.ctx, self.k, self.canonical = ctx, k, canonical
        self.h = _handle or lib().rg_count_begin(ctx.h, k, int(canonical), est_distinct, mode)
        if not self.h:
            raise _err("rg_count_begin")
        if shard is not None:
            _check(lib().rg_table_set_shard(self.h, shard[0], shard[1]), "rg_table_set_shard")

    # -- counting ------------------------------------------------------------
    def push_stream(self, stream: np.ndarray):
        """Host bytes; reads separated by any non-ACGT byte."""
        s = _u8(stream)
        _check(lib().rg_count_push(self.h, _ptr(s), s.size), "rg_count_push")

    def push_reads(self, seq: np.ndarray, offs: np.ndarray):
        s = _u8(seq)
        o = np.ascontiguousarray(offs, dtype=np.uint64)
        _check(lib().rg_count_push_reads(self.h, _ptr(s), _ptr(o), o.size - 1), "rg_count_push_reads")

    def push_dev(self, dev_ptr, n_bytes: int, read_stride: int = 0):
        _check(lib().rg_count_push_dev(self.h, _ptr(dev_ptr), n_bytes, read_stride), "rg_count_push_dev")

    def push_keys_dev(self, dev_ptr, n_keys: int):
        _check(lib().rg_count_push_keys_dev(self.h, _ptr(dev_ptr), n_keys), "rg_count_push_keys_dev")

    def end(self, trim: bool = True):
        _check(lib().rg_count_end(self.h), "rg_count_end")
        if trim:
            _check(lib().rg_table_trim(self.h), "rg_table_trim")
        return self

    def clear(self):
        _check(lib().rg_table_clear(self.h), "rg_table_clear")

    @classmethod
    def count(cls, ctx, seq, offs, k, canonical=True, est_distinct=0, mode=MODE_AUTO):
        """jellyfish count over (seq, offs) reads; grows and retries on RG_ERR_FULL."""
        est = est_distinct or max(1 << 16, int(_u8(seq).size) // 4)
        for _ in range(6):
            t = cls(ctx, k, canonical, est, mode)
            try:
                t.push_reads(seq, offs)
                return t.end()
            except RufusGpuError as e:
                t.free()
                if "full" not in str(e):
                    raise
                est *= 4
        raise RufusGpuError("count: table kept overflowing")

    @classmethod
    def from_pairs(cls, ctx, k, keys, counts, canonical=True):
        keys = np.ascontiguousarray(keys, np.uint64)
        counts = np.ascontiguousarray(counts, np.uint32)
        h = lib().rg_table_from_pairs(ctx.h, k, int(canonical), _ptr(keys), _ptr(counts), keys.size)
        if not h:
            raise _err("rg_table_from_pairs")
        return cls(ctx, k, canonical, _handle=h)

    def geometry(self):
        """tiles of this shard, hash-prefix groups over all shards, minimizer length, overflow slots"""
        nt, gg, m, ov = C.c_uint32(), C.c_uint32(), C.c_int(), C.c_uint64()
        _check(lib().rg_table_geometry(self.h, C.byref(nt), C.byref(gg), C.byref(m), C.byref(ov)), "rg_table_geometry")
        return {"n_tiles": nt.value, "groups": gg.value, "m": m.value, "ovf_slots": ov.value}

    # -- reading -------------------------------------------------------------
    def stats(self):
        d, o, s = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(lib().rg_table_stats(self.h, C.byref(d), C.byref(o), C.byref(s)), "rg_table_stats")
        return {"distinct": d.value, "occurrences": o.value, "slots": s.value}

    def histo(self, min_count: int = 1, low: int = 1, high: int = 10000) -> np.ndarray:
        out = np.zeros(lib().rg_histo_rows(low, high), dtype=np.uint64)
        _check(lib().rg_histo(self.h, min_count, low, high, _ptr(out)), "rg_histo")
        return out

    def dump(self, min_count: int = 1, max_count: int = 0xFFFFFFFF, sort: bool = True):
        cap = max(1024, self.stats()["distinct"])
        keys, counts = np.empty(cap, np.uint64), np.empty(cap, np.uint32)
        need = C.c_uint64()
        n = lib().rg_dump(self.h, min_count, max_count, _ptr(keys), _ptr(counts), cap, C.byref(need))
        if n < 0:
            raise _err("rg_dump")
        keys, counts = keys[:n], counts[:n]
        if sort:
            o = np.argsort(keys, kind="stable")
            keys, counts = keys[o], counts[o]
        return keys.copy(), counts.copy()

    def write_jhash(self, path: str, min_count: int = 1, max_count: int = 0, size: int = 1 << 33,
                    counter_len: int = 4, cmdline=()) -> int:
        """the .Jhash a `jellyfish count -L min_count -s size -o path` would leave (byte-compatible)"""
        argv = (C.c_char_p * max(len(cmdline), 1))(*[a.encode() for a in cmdline])
        n = lib().rg_table_write_jhash(self.h, path.encode(), min_count, max_count, size, counter_len, argv,
                                       len(cmdline))
        if n < 0:
            raise _err("rg_table_write_jhash")
        return int(n)

    @classmethod
    def read_jhash(cls, ctx, path: str):
        """load a .Jhash written by us or by real Jellyfish; returns (table, JfInfo)"""
        info = JfInfo()
        h = lib().rg_table_read_jhash(ctx.h, path.encode(), C.byref(info))
        if not h:
            raise _err("rg_table_read_jhash")
        return cls(ctx, info.k, bool(info.canonical), _handle=h), info

    def query(self, keys, min_count: int = 1) -> np.ndarray:
        keys = np.ascontiguousarray(keys, np.uint64)
        out = np.zeros(keys.size, np.uint32)
        _check(lib().rg_query(self.h, _ptr(keys), keys.size, min_count, _ptr(out)), "rg_query")
        return out

    def free(self):
        if self.h:
            lib().rg_table_free(self.h)
            self.h = None


def _handles(tabs):
    return (C.c_void_p * max(len(tabs), 1))(*[t.h for t in tabs])


def merge_unique(tabs, mins, floor: int = 5, sort: bool = True):
    """modified `jellyfish merge`: keys in exactly one table (count >= mins[i]) with count >= floor."""
    mins = np.ascontiguousarray(mins, np.uint64)
    cap = 1 << 16
    while True:
        keys, counts, which = np.empty(cap, np.uint64), np.empty(cap, np.uint32), np.empty(cap, np.int32)
        need = C.c_uint64()
        n = lib().rg_merge_unique(_handles(tabs), _ptr(mins), len(tabs), floor, _ptr(keys), _ptr(counts), _ptr(which),
                                  cap, C.byref(need))
        if n == ERR_SMALL:
            cap = int(need.value) + 1024
            continue
        if n < 0:
            raise _err("rg_merge_unique")
        keys, counts, which = keys[:n], counts[:n], which[:n]
        if sort:
            o = np.argsort(keys, kind="stable")
            keys, counts, which = keys[o], counts[o], which[o]
        return keys.copy(), counts.copy(), which.copy()


def subtract(subj, others, subj_min: int = 2, other_min=None, floor: int = 5, min_cov: int = 5, max_cov: int = 1200,
             sort: bool = True):
    """HashList = subject minus controls/excludes at the cut-offs (runRufus.sh:925-926)."""
    om = np.ascontiguousarray(other_min if other_min is not None else [2] * len(others), np.uint64)
    cap = 1 << 16
    while True:
        keys, counts = np.empty(cap, np.uint64), np.empty(cap, np.uint32)
        need = C.c_uint64()
        n = lib().rg_subtract(subj.h, subj_min, _handles(others), _ptr(om), len(others), floor, min_cov, max_cov,
                              _ptr(keys), _ptr(counts), cap, C.byref(need))
        if n == ERR_SMALL:
            cap = int(need.value) + 1024
            continue
        if n < 0:
            raise _err("rg_subtract")
        keys, counts = keys[:n], counts[:n]
        if sort:
            o = np.argsort(keys, kind="stable")
            keys, counts = keys[o], counts[o]
        return keys.copy(), counts.copy()


def build_select(subj, controls, min_subject: int, max_subject: int, max_control_sum: int, sort: bool = True):
    """RUFUS.Build (legacy): subject k-mers with min <= c_S <= max whose summed control count is <= max_control_sum.
    Returns (keys, subject counts, control sums)."""
    cap = 1 << 16
    while True:
        keys, counts, sums = np.empty(cap, np.uint64), np.empty(cap, np.uint32), np.empty(cap, np.uint32)
        need = C.c_uint64()
        n = lib().rg_build_select(subj.h, min_subject, max_subject, _handles(controls), len(controls), max_control_sum,
                                  _ptr(keys), _ptr(counts), _ptr(sums), cap, C.byref(need))
        if n == ERR_SMALL:
            cap = int(need.value) + 1024
            continue
        if n < 0:
            raise _err("rg_build_select")
        keys, counts, sums = keys[:n], counts[:n], sums[:n]
        if sort:
            o = np.argsort(keys, kind="stable")
            keys, counts, sums = keys[o], counts[o], sums[o]
        return keys.copy(), counts.copy(), sums.copy()


class MutantSet:
    """The `Mutations` hash of RUFUS.Filter (k-mer + reverse complement, RUFUS packing), in HBM."""

    def __init__(self, ctx: Context, handle, k: int):
        if not handle:
            raise _err("rg_set")
        self.ctx, self.h, self.k = ctx, handle, k

    @classmethod
    def from_hashlist_text(cls, ctx, text: bytes, k: int, paired: bool = True):
        return cls(ctx, lib().rg_set_from_hashlist_text(ctx.h, text, len(text), k, int(paired)), k)

    @classmethod
    def from_jf_keys(cls, ctx, keys, k: int):
        keys = np.ascontiguousarray(keys, np.uint64)
        return cls(ctx, lib().rg_set_from_jf_keys(ctx.h, _ptr(keys), keys.size, k), k)

    @classmethod
    def from_rufus_keys(cls, ctx, vals, k: int):
        vals = np.ascontiguousarray(vals, np.uint64)
        return cls(ctx, lib().rg_set_from_rufus_keys(ctx.h, _ptr(vals), vals.size, k), k)

    def __len__(self):
        return int(lib().rg_set_size(self.h))

    def filter_pairs(self, s1, q1, o1, s2, q2, o2, minq: int = 15, thresh: int = 1) -> np.ndarray:
        a = [_u8(x) for x in (s1, q1, s2, q2)]
        o1 = np.ascontiguousarray(o1, np.uint64)
        o2 = np.ascontiguousarray(o2, np.uint64)
        keep = np.zeros(o1.size - 1, np.uint8)
        _check(lib().rg_filter_pairs(self.h, _ptr(a[0]), _ptr(a[1]), _ptr(o1), _ptr(a[2]), _ptr(a[3]), _ptr(o2),
                                     o1.size - 1, minq, thresh, _ptr(keep)), "rg_filter_pairs")
        return keep

    def filter_single(self, s, q, o, minq: int = 15) -> np.ndarray:
        s, q = _u8(s), _u8(q)
        o = np.ascontiguousarray(o, np.uint64)
        hits = np.zeros(o.size - 1, np.int32)
        _check(lib().rg_filter_single(self.h, _ptr(s), _ptr(q), _ptr(o), o.size - 1, minq, _ptr(hits)),
               "rg_filter_single")
        return hits

    def annotate(self, s, q, o, minq: int = 3) -> np.ndarray:
        """AnnotateOverlap: per-base depth of mutant k-mer windows (the .hashcount.fastq quality line)"""
        s, q = _u8(s), _u8(q)
        o = np.ascontiguousarray(o, np.uint64)
        depth = np.zeros(int(o[-1]), np.uint8)
        _check(lib().rg_annotate(self.h, _ptr(s), _ptr(q), _ptr(o), o.size - 1, minq, _ptr(depth)), "rg_annotate")
        return depth

    def free(self):
        if self.h:
            lib().rg_set_free(self.h)
            self.h = None
