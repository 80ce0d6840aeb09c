"""ctypes wrapper of the oracle's C/OpenMP twin (TEST INFRASTRUCTURE ONLY).

Consumes the same packed stream layout as the GPU engine so that bench.py can
time "the CPU restatement of the Jellyfish path" on identical inputs
(SURVEY §8d CPU baseline plan).  Checked equal to the numpy oracle in
tests/test_oracle_c.py.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "libkdf_oracle.so")
_vp, _u64, _u32, _i = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
_lib = None

MODE_INSERT_COUNT, MODE_INSERT_ONLY, MODE_COUNT_IF_PRESENT, MODE_MARK_IF_PRESENT = 0, 1, 2, 3
U32_MAX = 0xFFFFFFFF


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB):
            raise RuntimeError("oracle C twin not built: run `bash oracle/build.sh`")
        l = ctypes.CDLL(LIB)
        l.okdf_table_new.restype = _vp
        l.okdf_table_new.argtypes = [_u64]
        l.okdf_table_free.argtypes = [_vp]
        l.okdf_clear_plane.argtypes = [_vp, _i]
        l.okdf_is_full.restype = _i
        l.okdf_is_full.argtypes = [_vp]
        l.okdf_count_stream.restype = _u64
        l.okdf_count_stream.argtypes = [_vp, _vp, _vp, _u64, _i, _i, _i, _u32, _i]
        l.okdf_update_keys.argtypes = [_vp, _vp, _vp, _u64, _i, _i, _u32, _i]
        l.okdf_threshold.restype = _u64
        l.okdf_threshold.argtypes = [_vp, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _u64]
        l.okdf_scan_reads.restype = _u64
        l.okdf_scan_reads.argtypes = [_vp, _vp, _vp, _u64, _vp, _vp, _u64, _i, _vp, _vp, _i]
        l.okdf_max_threads.restype = _i
        _lib = l
    return _lib


def _p(a):
    return a.ctypes.data_as(_vp) if a is not None else None


def max_threads():
    """Host threads the timed CPU legs use: every core this process may run on.  NOT
    ``omp_get_max_threads()``: torchrun exports OMP_NUM_THREADS=1 to its workers, which
    made the reference arm single-threaded (and time out) at N > 1 in round 1; the C
    functions take their thread count as an argument (``num_threads(threads)``)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class Table:
    def __init__(self, capacity):
        self.h = lib().okdf_table_new(int(capacity))
        if not self.h:
            raise MemoryError("okdf_table_new(%d)" % capacity)
        self.capacity = int(capacity)

    def close(self):
        if self.h:
            lib().okdf_table_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def count_stream(self, codes, valid, n_bases, k, mode=0, plane=0, arg=1, threads=1):
        codes = np.ascontiguousarray(codes, dtype=np.uint64)
        valid = np.ascontiguousarray(valid, dtype=np.uint32)
        n = lib().okdf_count_stream(self.h, _p(codes), _p(valid), int(n_bases), k, mode, plane,
                                    arg, threads)
        if lib().okdf_is_full(self.h):
            raise RuntimeError("oracle table full")
        return int(n)

    def update_keys(self, lo, hi, mode=1, plane=0, arg=0, threads=1):
        lo = np.ascontiguousarray(lo, dtype=np.uint64)
        hi = np.ascontiguousarray(hi, dtype=np.uint64) if hi is not None else None
        lib().okdf_update_keys(self.h, _p(lo), _p(hi), lo.shape[0], mode, plane, arg, threads)
        if lib().okdf_is_full(self.h):
            raise RuntimeError("oracle table full")

    def clear_plane(self, plane):
        lib().okdf_clear_plane(self.h, plane)

    def threshold(self, min0=0, max0=U32_MAX, min1=0, max1=U32_MAX, want=True):
        n = int(lib().okdf_threshold(self.h, min0, max0, min1, max1, None, None, None, None, 0))
        if not want:
            return n
        lo = np.zeros(max(n, 1), dtype=np.uint64)
        hi = np.zeros(max(n, 1), dtype=np.uint64)
        p0 = np.zeros(max(n, 1), dtype=np.uint32)
        p1 = np.zeros(max(n, 1), dtype=np.uint32)
        lib().okdf_threshold(self.h, min0, max0, min1, max1, _p(lo), _p(hi), _p(p0), _p(p1), n)
        return n, lo[:n], hi[:n], p0[:n], p1[:n]

    def scan_reads(self, codes, valid, n_bases, read_starts, read_lens, k, threads=1):
        codes = np.ascontiguousarray(codes, dtype=np.uint64)
        valid = np.ascontiguousarray(valid, dtype=np.uint32)
        rs = np.ascontiguousarray(read_starts, dtype=np.uint64)
        rl = np.ascontiguousarray(read_lens, dtype=np.uint32)
        nd = np.zeros(max(rs.shape[0], 1), dtype=np.uint32)
        nh = np.zeros(max(rs.shape[0], 1), dtype=np.uint32)
        nwin = lib().okdf_scan_reads(self.h, _p(codes), _p(valid), int(n_bases), _p(rs), _p(rl),
                                     rs.shape[0], k, _p(nd), _p(nh), threads)
        return nd[:rs.shape[0]], nh[:rs.shape[0]], int(nwin)


def discovery_chain(child, mother, father, ref, k, min_child_count=3, parent_max_count=0,
                    threads=1, child_capacity=None):
    """The k-mer part of the discovery path on packed streams.

    Each of child/mother/father/ref is ``(codes, valid, n_bases[, read_starts, read_lens])``.
    Returns dict with stage sizes, PU keys, per-read (nd, nh) and the number of
    k-mer instances processed (the unit of BASELINE.json's metric)."""
    units = 0
    cap = child_capacity or max(2 * int(child[2]), 1024)
    t = Table(cap)
    units += t.count_stream(child[0], child[1], child[2], k, MODE_INSERT_COUNT, 0, 1, threads)
    units += t.count_stream(ref[0], ref[1], ref[2], k, MODE_MARK_IF_PRESENT, 1, 1, threads)
    n_cand = t.threshold(min0=min_child_count, want=False)
    n_nonref, lo, hi, _a, _b = t.threshold(min0=min_child_count, max1=0)
    t.close()
    out = {"candidates": n_cand, "non_ref": n_nonref, "after_mother": 0, "proband_unique": 0,
           "pu_lo": np.zeros(0, np.uint64), "pu_hi": np.zeros(0, np.uint64), "nd": None, "nh": None}
    if n_nonref:
        tm = Table(max(2 * n_nonref, 1024))
        tm.update_keys(lo, hi, MODE_INSERT_ONLY, 0, 0, threads)
        units += tm.count_stream(mother[0], mother[1], mother[2], k, MODE_COUNT_IF_PRESENT, 0, 1, threads)
        n_am, lo, hi, _a, _b = tm.threshold(max0=parent_max_count)
        tm.close()
        out["after_mother"] = n_am
        if n_am:
            tf = Table(max(2 * n_am, 1024))
            tf.update_keys(lo, hi, MODE_INSERT_ONLY, 0, 0, threads)
            units += tf.count_stream(father[0], father[1], father[2], k, MODE_COUNT_IF_PRESENT, 0, 1, threads)
            n_pu, lo, hi, _a, _b = tf.threshold(max0=parent_max_count)
            tf.close()
            out.update({"proband_unique": n_pu, "pu_lo": lo, "pu_hi": hi})
            if n_pu and len(child) >= 5:
                tp = Table(max(2 * n_pu, 1024))
                tp.update_keys(lo, hi, MODE_INSERT_ONLY, 0, 0, threads)
                nd, nh, nwin = tp.scan_reads(child[0], child[1], child[2], child[3], child[4], k, threads)
                tp.close()
                units += nwin
                out.update({"nd": nd, "nh": nh})
    out["units"] = units
    return out
