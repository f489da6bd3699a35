"""oracle/oracle.py — TEST INFRASTRUCTURE: ctypes loaders for the two checkers.

* ``port()``  -> oracle/_build/liboracle_port.so   (plain-C restatement, oracle_port.c)
* ``ref()``   -> oracle/_ref/libdashing_ref_*.so   (ref_driver.cpp compiled against the unmodified
                                                     reference headers; built where /root/reference exists)

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import this module.  The product package ``dashing_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from functools import lru_cache

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "liboracle_port.so")
REF_DIR = os.path.join(HERE, "_ref")

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)


NEIGHBOR_DTYPE = np.dtype([("value", np.float32), ("index", np.uint32)])   # validx_t, src/sketch_and_cmp.h:605


def _ptr(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def build(port: bool = True, ref: bool = True) -> None:
    """(Re)build the checkers.  The reference driver is only rebuilt where /root/reference exists."""
    if port:
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    if ref and os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _cpu_flags() -> set:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


@lru_cache(maxsize=None)
def usable_cores() -> int:
    """Affinity mask capped by the cgroup CPU quota (the GPU boxes expose 128 logical CPUs but grant 16 CPUs of time;
    running 128 OpenMP threads against that quota is ~10x slower than running 16)."""
    import math
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            n = min(n, max(1, math.ceil(int(q) / int(per))))
    except Exception:
        pass
    return n


def ref_available() -> bool:
    return any(os.path.exists(os.path.join(REF_DIR, n)) for n in ("libdashing_ref_avx2.so", "libdashing_ref_avx512.so"))


@lru_cache(maxsize=None)
def _load_ref():
    flags = _cpu_flags()
    cands = []
    if {"avx512f", "avx512bw", "avx512dq", "avx512vl", "avx512cd"} <= flags:
        cands.append("libdashing_ref_avx512.so")
    cands.append("libdashing_ref_avx2.so")
    for n in cands:
        path = os.path.join(REF_DIR, n)
        if os.path.exists(path):
            return C.CDLL(path), n
    raise FileNotFoundError("oracle/_ref is not built (run `make -C oracle ref` where /root/reference exists)")


@lru_cache(maxsize=None)
def _load_port():
    if not os.path.exists(PORT_SO):
        build(port=True, ref=False)
    return C.CDLL(PORT_SO)


def pack_records(records):
    """records: list of bytes -> (bases uint8[], offsets uint64[nrec+1])"""
    offs = np.zeros(len(records) + 1, dtype=np.uint64)
    if records:
        offs[1:] = np.cumsum([len(r) for r in records], dtype=np.uint64)
    bases = np.frombuffer(b"".join(records), dtype=np.uint8).copy() if records else np.zeros(0, np.uint8)
    if bases.size == 0:
        bases = np.zeros(1, np.uint8)
    return bases, offs


class _Common:
    """Shared python-side conveniences; subclasses bind the C symbols."""

    def sketch(self, records, k, p, canon=True):
        bases, offs = pack_records(records)
        regs = np.zeros(1 << p, dtype=np.uint8)
        self._sketch(bases, offs, len(records), k, p, int(canon), regs)
        return regs


class Port(_Common):
    kind = "port"

    def __init__(self):
        l = self.l = _load_port()
        l.orc_wang.restype = C.c_uint64
        l.orc_wang.argtypes = [C.c_uint64]
        l.orc_canonical.restype = C.c_uint64
        l.orc_canonical.argtypes = [C.c_uint64, C.c_int]
        l.orc_kmers.restype = C.c_uint64
        l.orc_kmers.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int, u64p, C.c_uint64]
        l.orc_sketch.argtypes = [u8p, u64p, C.c_uint64, C.c_int, C.c_int, C.c_int, u8p]
        l.orc_histogram.argtypes = [u8p, C.c_int, u32p]
        l.orc_mle.restype = C.c_double
        l.orc_mle.argtypes = [u32p, C.c_int, C.c_int]
        l.orc_estimate.restype = C.c_double
        l.orc_estimate.argtypes = [u32p, C.c_int, C.c_int]
        l.orc_cardinality.restype = C.c_double
        l.orc_cardinality.argtypes = [u8p, C.c_int, C.c_int]
        l.orc_jaccard.restype = C.c_double
        l.orc_jaccard.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
        l.orc_union_size.restype = C.c_double
        l.orc_union_size.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
        l.orc_triple.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, f64p]
        l.orc_result_cmp.restype = C.c_float
        l.orc_result_cmp.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
        l.orc_dist_rows.argtypes = [u8p, C.c_uint64] + [C.c_int] * 6 + [C.c_uint64, C.c_uint64, f32p]
        l.orc_dist_rect.argtypes = [u8p, C.c_uint64, u8p, C.c_uint64] + [C.c_int] * 5 + [f32p]
        l.orc_knn.argtypes = [u8p, C.c_uint64] + [C.c_int] * 5 + [C.c_uint64, C.c_uint32, C.c_void_p]
        l.orc_hll_payload.restype = C.c_uint64
        l.orc_hll_payload.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_double, u8p, C.c_uint64]
        l.orc_union.argtypes = [u8p, C.c_uint64, C.c_int, u8p]
        l.orc_compress.argtypes = [u8p, C.c_int, C.c_int, u8p]

    def union(self, regs2d, p):
        regs2d = np.ascontiguousarray(regs2d, dtype=np.uint8).reshape(-1, 1 << p)
        out = np.zeros(1 << p, dtype=np.uint8)
        self.l.orc_union(_ptr(regs2d, u8p), regs2d.shape[0], p, _ptr(out, u8p))
        return out

    def compress(self, regs, p, new_p):
        out = np.zeros(1 << min(new_p, p), dtype=np.uint8)
        if self.l.orc_compress(_ptr(np.ascontiguousarray(regs, dtype=np.uint8), u8p), p, new_p, _ptr(out, u8p)):
            raise ValueError("Can't compress to a larger size")
        return out

    def wang(self, x):
        return int(self.l.orc_wang(C.c_uint64(x & 0xFFFFFFFFFFFFFFFF)))

    def kmers(self, s: bytes, k, canon=True):
        cap = max(len(s), 1)
        out = np.zeros(cap, dtype=np.uint64)
        n = self.l.orc_kmers(s, len(s), k, int(canon), _ptr(out, u64p), cap)
        return out[:n].copy()

    def _sketch(self, bases, offs, nrec, k, p, canon, regs):
        self.l.orc_sketch(_ptr(bases, u8p), _ptr(offs, u64p), nrec, k, p, canon, _ptr(regs, u8p))

    def histogram(self, regs, p):
        c = np.zeros(64, dtype=np.uint32)
        self.l.orc_histogram(_ptr(np.ascontiguousarray(regs), u8p), p, _ptr(c, u32p))
        return c

    def mle(self, c64, p, q=None):
        c = np.ascontiguousarray(c64, dtype=np.uint32)
        return float(self.l.orc_mle(_ptr(c, u32p), p, 64 - p if q is None else q))

    def estimate(self, c64, p, estim):
        c = np.ascontiguousarray(c64, dtype=np.uint32)
        return float(self.l.orc_estimate(_ptr(c, u32p), p, estim))

    def cardinality(self, regs, p, estim=2):
        return float(self.l.orc_cardinality(_ptr(np.ascontiguousarray(regs), u8p), p, estim))

    def cardinalities(self, regs2d, p, estim=2):
        return np.array([self.cardinality(r, p, estim) for r in regs2d], dtype=np.float64)

    def pair(self, lhs, rhs, p, estim=2, jestim=2, rtype=1, k=31):
        cl, cr = self.cardinality(lhs, p, estim), self.cardinality(rhs, p, estim)
        return float(self.l.orc_result_cmp(_ptr(np.ascontiguousarray(lhs), u8p), _ptr(np.ascontiguousarray(rhs), u8p),
                                           p, estim, jestim, rtype, k, cl, cr))

    def jaccard(self, lhs, rhs, p, estim=2, jestim=2):
        cl, cr = self.cardinality(lhs, p, estim), self.cardinality(rhs, p, estim)
        return float(self.l.orc_jaccard(_ptr(np.ascontiguousarray(lhs), u8p), _ptr(np.ascontiguousarray(rhs), u8p),
                                        p, estim, jestim, cl, cr))

    def triple(self, lhs, rhs, p, estim=2, jestim=2):
        cl, cr = self.cardinality(lhs, p, estim), self.cardinality(rhs, p, estim)
        out = np.zeros(3)
        self.l.orc_triple(_ptr(np.ascontiguousarray(lhs), u8p), _ptr(np.ascontiguousarray(rhs), u8p), p, estim, jestim,
                          cl, cr, _ptr(out, f64p))
        return out

    def dist_rows(self, regs2d, p, k=31, estim=2, jestim=2, rtype=1, order=0, row_begin=0, row_end=None, nthreads=0):
        regs2d = np.ascontiguousarray(regs2d, dtype=np.uint8)
        n = regs2d.shape[0]
        out = np.zeros(n * (n - 1) // 2, dtype=np.float32)
        self.l.orc_dist_rows(_ptr(regs2d, u8p), n, p, k, estim, jestim, rtype, order, row_begin,
                             n if row_end is None else row_end, _ptr(out, f32p))
        return out

    def dist_symmetric(self, regs2d, p, **kw):
        return self.dist_rows(regs2d, p, **kw)

    def dist_rect(self, refs, qrys, p, k=31, estim=2, jestim=2, rtype=1, nthreads=0):
        refs = np.ascontiguousarray(refs, dtype=np.uint8)
        qrys = np.ascontiguousarray(qrys, dtype=np.uint8)
        out = np.zeros((qrys.shape[0], refs.shape[0]), dtype=np.float32)
        self.l.orc_dist_rect(_ptr(refs, u8p), refs.shape[0], _ptr(qrys, u8p), qrys.shape[0], p, k, estim, jestim, rtype,
                             _ptr(out, f32p))
        return out

    def knn(self, regs2d, p, nn, k=31, estim=2, jestim=2, rtype=1, nq=0, nthreads=1):
        """[rows][nn] (value, index), rows = nq or n; the last nq sketches are the queries when nq > 0."""
        regs2d = np.ascontiguousarray(regs2d, dtype=np.uint8)
        n = regs2d.shape[0]
        out = np.zeros((nq if nq else n, nn), dtype=NEIGHBOR_DTYPE)
        if self.l.orc_knn(_ptr(regs2d, u8p), n, p, k, estim, jestim, rtype, nq, nn, C.c_void_p(out.ctypes.data)):
            raise RuntimeError("orc_knn failed")
        return out

    def hll_payload(self, regs, p, estim=2, jestim=2, value=-1.0):
        out = np.zeros(28 + (1 << p), dtype=np.uint8)
        n = self.l.orc_hll_payload(_ptr(np.ascontiguousarray(regs), u8p), p, estim, jestim, value, _ptr(out, u8p), out.size)
        assert n == out.size
        return out.tobytes()


class Ref(_Common):
    kind = "reference"

    def __init__(self, lib=None, libname=None):
        l, self.libname = (lib, libname) if lib is not None else _load_ref()
        self.l = l
        l.dref_simd_tier.restype = C.c_int
        l.dref_max_threads.restype = C.c_int
        l.dref_wang.restype = C.c_uint64
        l.dref_wang.argtypes = [C.c_uint64]
        l.dref_kmers.restype = C.c_uint64
        l.dref_kmers.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int, u64p, C.c_uint64]
        l.dref_sketch.argtypes = [u8p, u64p, C.c_uint64, C.c_int, C.c_int, C.c_int, u8p]
        l.dref_sketch_many.argtypes = [u8p, u64p, u64p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        l.dref_histogram.argtypes = [u8p, C.c_int, u32p]
        l.dref_mle.restype = C.c_double
        l.dref_mle.argtypes = [u32p, C.c_int, C.c_int]
        l.dref_estimate_from_counts.restype = C.c_double
        l.dref_estimate_from_counts.argtypes = [u32p, C.c_int, C.c_int]
        l.dref_cardinality.restype = C.c_double
        l.dref_cardinality.argtypes = [u8p, C.c_int, C.c_int]
        l.dref_cardinalities.argtypes = [u8p, C.c_uint64, C.c_int, C.c_int, f64p]
        l.dref_pair.restype = C.c_float
        l.dref_pair.argtypes = [u8p, u8p] + [C.c_int] * 5
        l.dref_jaccard.restype = C.c_double
        l.dref_jaccard.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int]
        l.dref_union_size.restype = C.c_double
        l.dref_union_size.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int]
        l.dref_triple.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, f64p]
        l.dref_dist_rows.argtypes = [u8p, C.c_uint64] + [C.c_int] * 6 + [C.c_uint64, C.c_uint64, C.c_int, f32p]
        l.dref_dist_rect.argtypes = [u8p, C.c_uint64, u8p, C.c_uint64] + [C.c_int] * 6 + [f32p]
        l.dref_set_create.restype = C.c_void_p
        l.dref_set_create.argtypes = [u8p, C.c_uint64, C.c_int, C.c_int, C.c_int]
        l.dref_set_free.argtypes = [C.c_void_p]
        l.dref_set_dist_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_int, f32p]
        l.dref_knn.argtypes = [u8p, C.c_uint64] + [C.c_int] * 5 + [C.c_uint64, C.c_uint32, C.c_int, C.c_void_p]
        l.dref_set_nneighbors.argtypes = [C.c_uint32]
        l.dref_cli_dist.argtypes = [C.c_int, C.POINTER(C.c_char_p)] + [C.c_int] * 10 + [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p]
        l.dref_cli_sketch.argtypes = [C.c_int, C.POINTER(C.c_char_p)] + [C.c_int] * 4 + [C.c_char_p, C.c_char_p, C.c_int]
        l.dref_hll_write.argtypes = [C.c_char_p, u8p, C.c_int, C.c_int, C.c_int, C.c_int]
        l.dref_hll_read.argtypes = [C.c_char_p, u8p, C.c_uint64, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                    C.POINTER(C.c_int), f64p]
        l.dref_make_fname.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p,
                                      C.c_char_p, C.c_uint64]

        l.dref_compress.argtypes = [u8p, C.c_int, C.c_int, u8p]
        l.dref_union.argtypes = [u8p, C.c_uint64, C.c_int, u8p]
        l.dref_union_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
        l.dref_hll_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
        l.dref_fold.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        l.dref_view.argtypes = [C.c_char_p, C.c_char_p]
        l.dref_cli_sketch_container.argtypes = [C.c_int, C.POINTER(C.c_char_p)] + [C.c_int] * 4 + [C.c_char_p]
        l.dref_cli_sketch_by_seq.argtypes = [C.c_char_p, C.c_char_p] + [C.c_int] * 5
        l.dref_cli_dist_by_seq.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p] + [C.c_int] * 6
        l.dref_cli_card.argtypes = [C.c_int, C.POINTER(C.c_char_p)] + [C.c_int] * 9 + [C.c_char_p]
        l.dref_cli_dist_defer.argtypes = [C.c_int, C.POINTER(C.c_char_p)] + [C.c_int] * 9 + [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p]

    # ---- SURVEY.md §8(f)3 -------------------------------------------------------------------------------------------
    def union(self, regs2d, p):
        regs2d = np.ascontiguousarray(regs2d, dtype=np.uint8).reshape(-1, 1 << p)
        out = np.zeros(1 << p, dtype=np.uint8)
        self.l.dref_union(_ptr(regs2d, u8p), regs2d.shape[0], p, _ptr(out, u8p))
        return out

    def compress(self, regs, p, new_p):
        out = np.zeros(1 << min(new_p, p), dtype=np.uint8)
        if self.l.dref_compress(_ptr(np.ascontiguousarray(regs, dtype=np.uint8), u8p), p, new_p, _ptr(out, u8p)):
            raise ValueError("Can't compress to a larger size")
        return out

    @staticmethod
    def _argv(args):
        return (C.c_char_p * (len(args) + 1))(*[os.fsencode(a) for a in args], None)

    def union_main(self, *args):
        """The reference's own `dashing union` main (src/union.cpp:60-108)."""
        a = ["union", *args]
        if self.l.dref_union_main(len(a), self._argv(a)):
            raise RuntimeError("dref_union_main failed")

    def hll_main(self, *args):
        """The reference's own `dashing hll` main (src/hllmain.cpp); prints to this process's stdout."""
        a = ["hll", *args]
        if self.l.dref_hll_main(len(a), self._argv(a)):
            raise RuntimeError("dref_hll_main failed")

    def fold(self, inp, out, destp=-1):
        if self.l.dref_fold(os.fsencode(inp), os.fsencode(out), destp):
            raise RuntimeError("dref_fold failed")

    def view(self, inp, out):
        if self.l.dref_view(os.fsencode(inp), os.fsencode(out)):
            raise RuntimeError("dref_view failed")

    def cli_sketch_container(self, paths, output_file, k=31, p=10, canon=True, nthreads=None):
        """`dashing sketch -o`; the reference indexes its worker sketches by path index here, so nthreads >= len(paths)."""
        arr = (C.c_char_p * len(paths))(*[os.fsencode(p_) for p_ in paths])
        if self.l.dref_cli_sketch_container(len(paths), arr, k, p, int(canon), nthreads or len(paths), os.fsencode(output_file)):
            raise RuntimeError("dref_cli_sketch_container failed")

    def cli_sketch_by_seq(self, inpath, outpath, k=31, p=10, canon=True, estim=2, jestim=2):
        if self.l.dref_cli_sketch_by_seq(os.fsencode(inpath), os.fsencode(outpath), k, p, int(canon), estim, jestim):
            raise RuntimeError("dref_cli_sketch_by_seq failed")

    def cli_dist_by_seq(self, namefile, datapath, outpath, k=31, estim=2, jestim=2, rtype=1, emit_fmt=0, nthreads=1):
        if self.l.dref_cli_dist_by_seq(os.fsencode(namefile), os.fsencode(datapath), os.fsencode(outpath), k, estim, jestim, rtype,
                                       emit_fmt, nthreads, 0):
            raise RuntimeError("dref_cli_dist_by_seq failed")

    def cli_card(self, paths, outpath, k=31, p=10, canon=True, estim=2, jestim=2, presketched=False, emit_binary=False,
                 use_scientific=False, nthreads=1):
        arr = (C.c_char_p * len(paths))(*[os.fsencode(p_) for p_ in paths])
        if self.l.dref_cli_card(len(paths), arr, k, p, int(canon), estim, jestim, int(presketched), int(emit_binary),
                                int(use_scientific), nthreads, os.fsencode(outpath)):
            raise RuntimeError("dref_cli_card failed")

    def cli_dist_defer(self, paths, sizes_path, dist_path, nq=0, k=31, p=10, canon=True, estim=2, jestim=2, rtype=1, emit_fmt=0,
                       nthreads=1, cache=False, prefix="", suffix=""):
        """`dashing dist --defer-hll` (dist_sketch_and_cmp<HyperLogLogHasher<>>)."""
        arr = (C.c_char_p * len(paths))(*[os.fsencode(p_) for p_ in paths])
        if self.l.dref_cli_dist_defer(len(paths), arr, nq, k, p, int(canon), estim, jestim, rtype, emit_fmt, nthreads,
                                      os.fsencode(sizes_path), os.fsencode(dist_path), int(cache), prefix.encode(), suffix.encode()):
            raise RuntimeError("dref_cli_dist_defer failed")

    def simd_tier(self):
        return int(self.l.dref_simd_tier())

    def max_threads(self):
        return int(self.l.dref_max_threads())

    def wang(self, x):
        return int(self.l.dref_wang(C.c_uint64(x & 0xFFFFFFFFFFFFFFFF)))

    def kmers(self, s: bytes, k, canon=True):
        cap = max(len(s), 1)
        out = np.zeros(cap, dtype=np.uint64)
        n = self.l.dref_kmers(s, len(s), k, int(canon), _ptr(out, u64p), cap)
        return out[:n].copy()

    def _sketch(self, bases, offs, nrec, k, p, canon, regs):
        self.l.dref_sketch(_ptr(bases, u8p), _ptr(offs, u64p), nrec, k, p, canon, _ptr(regs, u8p))

    def sketch_many(self, bases, offs, genome_rec_begin, k, p, canon=True, nthreads=0):
        ng = len(genome_rec_begin) - 1
        regs = np.zeros((ng, 1 << p), dtype=np.uint8)
        grb = np.ascontiguousarray(genome_rec_begin, dtype=np.uint64)
        self.l.dref_sketch_many(_ptr(bases, u8p), _ptr(offs, u64p), _ptr(grb, u64p), ng, k, p, int(canon), nthreads or usable_cores(),
                                _ptr(regs, u8p))
        return regs

    def histogram(self, regs, p):
        c = np.zeros(64, dtype=np.uint32)
        self.l.dref_histogram(_ptr(np.ascontiguousarray(regs), u8p), p, _ptr(c, u32p))
        return c

    def mle(self, c64, p, q=None):
        c = np.ascontiguousarray(c64, dtype=np.uint32)
        return float(self.l.dref_mle(_ptr(c, u32p), p, 64 - p if q is None else q))

    def estimate(self, c64, p, estim):
        c = np.ascontiguousarray(c64, dtype=np.uint32)
        return float(self.l.dref_estimate_from_counts(_ptr(c, u32p), p, estim))

    def cardinality(self, regs, p, estim=2):
        return float(self.l.dref_cardinality(_ptr(np.ascontiguousarray(regs), u8p), p, estim))

    def cardinalities(self, regs2d, p, estim=2):
        regs2d = np.ascontiguousarray(regs2d, dtype=np.uint8)
        out = np.zeros(regs2d.shape[0])
        self.l.dref_cardinalities(_ptr(regs2d, u8p), regs2d.shape[0], p, estim, _ptr(out, f64p))
        return out

    def pair(self, lhs, rhs, p, estim=2, jestim=2, rtype=1, k=31):
        return float(self.l.dref_pair(_ptr(np.ascontiguousarray(lhs), u8p), _ptr(np.ascontiguousarray(rhs), u8p),
                                      p, estim, jestim, rtype, k))

    def jaccard(self, lhs, rhs, p, estim=2, jestim=2):
        return float(self.l.dref_jaccard(_ptr(np.ascontiguousarray(lhs), u8p), _ptr(np.ascontiguousarray(rhs), u8p),
                                         p, estim, jestim))

    def union_size(self, lhs, rhs, p, estim=2, jestim=2):
        return float(self.l.dref_union_size(_ptr(np.ascontiguousarray(lhs), u8p), _ptr(np.ascontiguousarray(rhs), u8p),
                                            p, estim, jestim))

    def triple(self, lhs, rhs, p, estim=2, jestim=2):
        out = np.zeros(3)
        self.l.dref_triple(_ptr(np.ascontiguousarray(lhs), u8p), _ptr(np.ascontiguousarray(rhs), u8p), p, estim, jestim,
                           _ptr(out, f64p))
        return out

    def dist_rows(self, regs2d, p, k=31, estim=2, jestim=2, rtype=1, order=0, row_begin=0, row_end=None, nthreads=0,
                  out=None):
        regs2d = np.ascontiguousarray(regs2d, dtype=np.uint8)
        n = regs2d.shape[0]
        if out is None:
            out = np.zeros(n * (n - 1) // 2, dtype=np.float32)
        self.l.dref_dist_rows(_ptr(regs2d, u8p), n, p, k, estim, jestim, rtype, order, row_begin,
                              n if row_end is None else row_end, nthreads or usable_cores(), _ptr(out, f32p))
        return out

    def dist_symmetric(self, regs2d, p, **kw):
        return self.dist_rows(regs2d, p, **kw)

    # prepared set: build the n hll_t objects once, then time row ranges of the pair loop only
    def set_create(self, regs2d, p, estim=2, jestim=2):
        regs2d = np.ascontiguousarray(regs2d, dtype=np.uint8)
        return (self.l.dref_set_create(_ptr(regs2d, u8p), regs2d.shape[0], p, estim, jestim), regs2d.shape[0])

    def set_dist_rows(self, hset, k, rtype, order, row_begin, row_end, nthreads=0):
        h, n = hset
        tri = lambda r: r * (2 * n - r - 1) // 2
        out = np.zeros(max(tri(min(row_end, n)) - tri(row_begin), 1), dtype=np.float32)
        self.l.dref_set_dist_rows(C.c_void_p(h), k, rtype, order, row_begin, row_end, nthreads or usable_cores(), _ptr(out, f32p))
        return out

    def set_free(self, hset):
        self.l.dref_set_free(C.c_void_p(hset[0]))

    def dist_rect(self, refs, qrys, p, k=31, estim=2, jestim=2, rtype=1, nthreads=0):
        refs = np.ascontiguousarray(refs, dtype=np.uint8)
        qrys = np.ascontiguousarray(qrys, dtype=np.uint8)
        out = np.zeros((qrys.shape[0], refs.shape[0]), dtype=np.float32)
        self.l.dref_dist_rect(_ptr(refs, u8p), refs.shape[0], _ptr(qrys, u8p), qrys.shape[0], p, k, estim, jestim, rtype,
                              nthreads or usable_cores(), _ptr(out, f32p))
        return out

    def knn(self, regs2d, p, nn, k=31, estim=2, jestim=2, rtype=1, nq=0, nthreads=1):
        """The reference's perform_nns; nthreads=1 is the deterministic order."""
        regs2d = np.ascontiguousarray(regs2d, dtype=np.uint8)
        n = regs2d.shape[0]
        out = np.zeros((nq if nq else n, nn), dtype=NEIGHBOR_DTYPE)
        self.l.dref_knn(_ptr(regs2d, u8p), n, p, k, estim, jestim, rtype, nq, nn, nthreads, C.c_void_p(out.ctypes.data))
        return out

    def hll_write(self, path, regs, p, estim=2, jestim=2, calculated=False):
        rc = self.l.dref_hll_write(os.fsencode(path), _ptr(np.ascontiguousarray(regs), u8p), p, estim, jestim,
                                   int(calculated))
        if rc:
            raise RuntimeError("dref_hll_write failed")

    def hll_read(self, path, max_p=24):
        regs = np.zeros(1 << max_p, dtype=np.uint8)
        p, e, j, v = C.c_int(), C.c_int(), C.c_int(), C.c_double()
        rc = self.l.dref_hll_read(os.fsencode(path), _ptr(regs, u8p), regs.size, C.byref(p), C.byref(e), C.byref(j),
                                  C.byref(v))
        if rc:
            raise RuntimeError("dref_hll_read failed")
        return regs[: 1 << p.value].copy(), p.value, e.value, j.value, v.value

    def cli_dist(self, paths, sizes_path, dist_path, nq=0, k=31, p=10, canon=True, estim=2, jestim=2, rtype=1, emit_fmt=0,
                 presketched=False, nthreads=1, cache=False, prefix="", suffix="", nneighbors=0):
        """The reference's dist_sketch_and_cmp<hll_t> (sizes file + distance output), paths in final order."""
        arr = (C.c_char_p * len(paths))(*[os.fsencode(p_) for p_ in paths])
        self.l.dref_set_nneighbors(nneighbors)
        if nneighbors:
            emit_fmt |= 8          # NEAREST_NEIGHBOR_TABLE, src/enums.h:32; src/distmain.cpp:106-109
        rc = self.l.dref_cli_dist(len(paths), arr, nq, k, p, int(canon), estim, jestim, rtype, emit_fmt, int(presketched), nthreads,
                                  os.fsencode(sizes_path), os.fsencode(dist_path), int(cache), prefix.encode(), suffix.encode())
        if rc:
            raise RuntimeError(f"dref_cli_dist failed ({rc})")

    def cli_sketch(self, paths, k=31, p=10, canon=True, nthreads=1, prefix="", suffix="", skip_cached=False):
        """The reference's sketch_core<hll_t>: one .hll per path."""
        arr = (C.c_char_p * len(paths))(*[os.fsencode(p_) for p_ in paths])
        rc = self.l.dref_cli_sketch(len(paths), arr, k, p, int(canon), nthreads, prefix.encode(), suffix.encode(), int(skip_cached))
        if rc:
            raise RuntimeError(f"dref_cli_sketch failed ({rc})")

    def make_fname(self, path, p, wsz, k, csz, spacing="", suffix="", prefix=""):
        buf = C.create_string_buffer(4096)
        rc = self.l.dref_make_fname(os.fsencode(path), p, wsz, k, csz, spacing.encode(), suffix.encode(), prefix.encode(),
                                    buf, 4096)
        if rc:
            raise RuntimeError("dref_make_fname failed")
        return buf.value.decode()


@lru_cache(maxsize=None)
def port() -> Port:
    return Port()


@lru_cache(maxsize=None)
def ref() -> Ref:
    return Ref()


PATCHED_SO = os.path.join(REF_DIR, "libdashing_ref_patched.so")


def patched_available() -> bool:
    return os.path.exists(PATCHED_SO)


@lru_cache(maxsize=None)
def ref_patched() -> Ref:
    """The SAME driver built against the reference headers with oracle/integration.patch applied (-DDASHING_B200): its
    sketch_core / dist_sketch_and_cmp / dist_loop / partdist_loop run their hot paths through libdashing_b200's C ABI.
    Needs a CUDA device for anything that sketches or compares (DASHING_GPU=0 in the environment switches it back to the
    reference's own code paths).  Test infrastructure for INTEGRATION.md, like everything else under oracle/."""
    if not patched_available():
        raise FileNotFoundError("oracle/_ref/libdashing_ref_patched.so is not built (run `make -C oracle patched` where /root/reference exists)")
    return Ref(C.CDLL(PATCHED_SO), "libdashing_ref_patched.so")


def best():
    """The strongest checker available: the real reference if built, else the pinned C port."""
    return ref() if ref_available() else port()
