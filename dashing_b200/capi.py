"""ctypes binding of include/dashing_b200.h (libdashing_b200.so).

Host buffers are numpy arrays; ``*_dev`` wrappers take raw device pointers (ints) and a CUDA stream
handle so callers (bench.py, the multi-GPU driver) can use torch tensors / streams for the plumbing.
There is no Python or CPU implementation behind these functions: if the shared library is missing
the import fails, and without a CUDA device every compute call raises ``Db200Error`` (DB200_ENODEV).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdashing_b200.so")

OK, EINVAL, EUNSUPPORTED, ENODEV, ECUDA, ENOMEM = range(6)
ORIGINAL, ERTL_IMPROVED, ERTL_MLE, ERTL_JOINT_MLE = 0, 1, 2, 3
MASH_DIST, JI, SIZES, FULL_MASH_DIST, FULL_CONTAINMENT_DIST, CONTAINMENT_INDEX, CONTAINMENT_DIST, \
    SYMMETRIC_CONTAINMENT_INDEX, SYMMETRIC_CONTAINMENT_DIST, UNION_SIZE = range(10)
ORDER_ROW_FIRST, ORDER_COL_FIRST = 0, 1
ALL_DEVICES = -1   # DB200_ALL_DEVICES: host-pointer entry points shard over every visible GPU


class Db200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libdashing_b200 error {code}: {msg}")
        self.code = code


class DistParams(C.Structure):
    _fields_ = [("p", C.c_int32), ("k", C.c_int32), ("estim", C.c_int32), ("jestim", C.c_int32),
                ("result_type", C.c_int32), ("order", C.c_int32),
                ("card", C.POINTER(C.c_double)), ("card_queries", C.POINTER(C.c_double))]


u8p, u64p, f32p, f64p, vp = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_float), C.POINTER(C.c_double), C.c_void_p
ROWS_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, f32p, C.c_uint64)

if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} is not built — run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(nvcc, sm_100a).  dashing_b200 has no fallback implementation.")
lib = C.CDLL(LIB_PATH)

_SIGS = {
    "db200_last_error": (C.c_char_p, []),
    "db200_version": (C.c_int, []),
    "db200_device_count": (C.c_int, []),
    "db200_warmup": (C.c_int, [C.c_int]),
    "db200_host_alloc": (C.c_int, [C.POINTER(vp), C.c_size_t]),
    "db200_host_free": (C.c_int, [vp]),
    "db200_sketcher_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.POINTER(vp)]),
    "db200_sketcher_add_record": (C.c_int, [vp, C.c_uint32, C.c_char_p, C.c_uint64]),
    "db200_sketcher_finish": (C.c_int, [vp, C.c_uint32, u8p]),
    "db200_sketcher_destroy": (C.c_int, [vp]),
    "db200_sketch_batch": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, vp, u64p, C.c_uint64, u64p, C.c_uint64, u8p]),
    "db200_sketch_fasta_batch": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, vp, u64p, u64p, C.c_uint64, u64p, C.c_uint64, u8p, u8p]),
    "db200_hostpack": (None, [vp, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint16)]),
    "db200_hostpack_isa": (C.c_char_p, []),
    "db200_pack_genomes": (C.c_int, [C.c_int, vp, u64p, C.c_uint64, u64p, C.c_uint64, C.c_int, C.POINTER(vp)]),
    "db200_repack_genomes": (C.c_int, [vp, vp, u64p, C.c_uint64, u64p, C.c_uint64, C.c_int]),
    "db200_packed_genomes_free": (C.c_int, [vp]),
    "db200_packed_genomes_stats": (C.c_int, [vp, u64p, u64p, u64p]),
    "db200_sketch_packed_dev": (C.c_int, [vp, C.c_int, C.c_int, vp, vp]),
    "db200_cardinalities": (C.c_int, [C.c_int, u8p, C.c_uint64, C.c_int, C.c_int, f64p]),
    "db200_union": (C.c_int, [C.c_int, u8p, C.c_uint64, C.c_int, u8p]),
    "db200_compress": (C.c_int, [C.c_int, u8p, C.c_uint64, C.c_int, C.c_int, u8p]),
    "db200_dist_symmetric": (C.c_int, [C.c_int, u8p, C.c_uint64, C.POINTER(DistParams), f32p]),
    "db200_dist_symmetric_rows": (C.c_int, [C.c_int, u8p, C.c_uint64, C.POINTER(DistParams), C.c_uint64, C.c_uint64, f32p]),
    "db200_dist_symmetric_stream": (C.c_int, [C.c_int, u8p, C.c_uint64, C.POINTER(DistParams), C.c_uint64, C.c_uint64, C.c_uint64, ROWS_CB, vp]),
    "db200_dist_rect": (C.c_int, [C.c_int, u8p, C.c_uint64, u8p, C.c_uint64, C.POINTER(DistParams), f32p]),
    "db200_dist_plan_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "db200_dist_plan_destroy": (C.c_int, [vp]),
    "db200_dist_plan_prepare_dev": (C.c_int, [vp, vp, C.c_uint64, C.c_int, C.c_int, vp]),
    "db200_dist_plan_begin_dev": (C.c_int, [vp, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "db200_dist_plan_add_rows_dev": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, vp]),
    "db200_dist_plan_finish_dev": (C.c_int, [vp]),
    "db200_dist_plan_run_symmetric_dev": (C.c_int, [vp, C.POINTER(DistParams), C.c_uint64, C.c_uint64, vp, vp]),
    "db200_dist_plan_run_rect_dev": (C.c_int, [vp, C.POINTER(DistParams), C.c_uint64, C.c_uint64, vp, vp]),
    "db200_dist_knn_symmetric": (C.c_int, [C.c_int, u8p, C.c_uint64, C.POINTER(DistParams), C.c_uint32, vp]),
    "db200_dist_knn_rect": (C.c_int, [C.c_int, u8p, C.c_uint64, u8p, C.c_uint64, C.POINTER(DistParams), C.c_uint32, vp]),
    "db200_dist_plan_run_knn_dev": (C.c_int, [vp, C.POINTER(DistParams), C.c_uint64, C.c_uint64, C.c_uint32, vp, vp]),
    "db200_dist_plan_run_knn_rows_dev": (C.c_int, [vp, C.POINTER(DistParams), C.c_uint64, C.c_uint64, C.c_uint32, vp, vp]),
    "db200_dist_plan_cardinalities_dev": (C.c_int, [vp, C.POINTER(vp)]),
    "db200_kernel_launches": (C.c_uint64, []),
    "db200_dist_plan_last_run_info": (C.c_int, [vp, u64p, u64p, C.POINTER(C.c_int)]),
}
for _name, (_res, _args) in _SIGS.items():
    _f = getattr(lib, _name)  # AttributeError here == the .so does not export what the header declares
    _f.restype, _f.argtypes = _res, _args

EXPORTS = tuple(_SIGS)


def _check(rc: int):
    if rc != OK:
        raise Db200Error(rc, (lib.db200_last_error() or b"").decode(errors="replace"))


def _np(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def device_count() -> int:
    return int(lib.db200_device_count())


def kernel_launches() -> int:
    return int(lib.db200_kernel_launches())


def dist_params(p, k=31, estim=ERTL_MLE, jestim=ERTL_MLE, result_type=JI, order=ORDER_ROW_FIRST, card=None, card_queries=None) -> DistParams:
    """card / card_queries: cached per-sketch cardinalities (db200_dist_params.card; float64 arrays, kept alive on the returned
    object), None = the library evaluates them from the registers."""
    prm = DistParams(p, k, estim, jestim, result_type, order, None, None)
    keep = []
    for name, arr in (("card", card), ("card_queries", card_queries)):
        if arr is not None:
            arr = np.ascontiguousarray(arr, dtype=np.float64)
            keep.append(arr)
            setattr(prm, name, arr.ctypes.data_as(f64p))
    prm._keep = keep
    return prm


def pinned_empty(nbytes: int) -> np.ndarray:
    """uint8 numpy view of page-locked host memory (never freed: meant for long-lived bench buffers)."""
    ptr = vp()
    _check(lib.db200_host_alloc(C.byref(ptr), nbytes))
    return np.ctypeslib.as_array(C.cast(ptr, u8p), shape=(max(nbytes, 1),))[:nbytes]


# ---- sketching ------------------------------------------------------------------------------
def records_layout(genomes):
    """genomes: list of genomes, each a list of records (bytes / uint8 arrays) or a single record.
    -> (bases uint8[T], rec_offsets uint64[nrec+1], genome_rec_begin uint64[ng+1])"""
    recs, grb = [], [0]
    for g in genomes:
        rs = g if isinstance(g, (list, tuple)) else [g]
        for r in rs:
            recs.append(np.frombuffer(r, dtype=np.uint8) if isinstance(r, (bytes, bytearray)) else np.asarray(r, dtype=np.uint8))
        grb.append(len(recs))
    offs = np.zeros(len(recs) + 1, dtype=np.uint64)
    if recs:
        offs[1:] = np.cumsum([r.size for r in recs], dtype=np.uint64)
    bases = np.concatenate(recs) if recs else np.zeros(0, np.uint8)
    if bases.size == 0:
        bases = np.zeros(1, np.uint8)
    return np.ascontiguousarray(bases), offs, np.asarray(grb, dtype=np.uint64)


def sketch_batch(bases, rec_offsets, genome_rec_begin, k, p, canon=True, device=0) -> np.ndarray:
    bases = _np(bases, np.uint8)
    offs = _np(rec_offsets, np.uint64)
    grb = _np(genome_rec_begin, np.uint64)
    ng = grb.size - 1
    out = np.zeros((ng, 1 << p), dtype=np.uint8)
    _check(lib.db200_sketch_batch(device, p, k, int(canon), bases.ctypes.data_as(vp), offs.ctypes.data_as(u64p), offs.size - 1,
                                  grb.ctypes.data_as(u64p), ng, out.ctypes.data_as(u8p)))
    return out


FASTA_ALIGN = 8192


def fasta_layout(genome_files):
    """genome_files: list of genomes, each a list of raw file contents (bytes).
    -> (text uint8[], file_off uint64[nf], file_len uint64[nf], genome_file_begin uint64[ng+1]) with every file on a
    DB200_FASTA_ALIGN boundary (gaps are filled with junk on purpose: the library must ignore them)."""
    offs, lens, gfb, parts, at = [], [], [0], [], 0
    for files in genome_files:
        for raw in files:
            offs.append(at)
            lens.append(len(raw))
            pad = (-len(raw)) % FASTA_ALIGN or (FASTA_ALIGN if len(raw) == 0 else 0)
            junk = (b">junk\nACGTTGCA\r\n@q\n+\n" * (pad // 16 + 1))[:pad]
            parts.append(bytes(raw) + junk)
            at += len(raw) + pad
        gfb.append(len(offs))
    text = np.frombuffer(b"".join(parts) + b"\0" * 64, dtype=np.uint8).copy()
    return text, np.asarray(offs, np.uint64), np.asarray(lens, np.uint64), np.asarray(gfb, np.uint64)


def sketch_fasta(genome_files, k, p, canon=True, device=0):
    """Raw FASTA text -> (registers uint8[ng][2^p], file_status uint8[nf]) through db200_sketch_fasta_batch."""
    text, offs, lens, gfb = fasta_layout(genome_files)
    ng = gfb.size - 1
    out = np.zeros((ng, 1 << p), dtype=np.uint8)
    status = np.zeros(max(offs.size, 1), dtype=np.uint8)
    _check(lib.db200_sketch_fasta_batch(device, p, k, int(canon), text.ctypes.data_as(vp), offs.ctypes.data_as(u64p), lens.ctypes.data_as(u64p),
                                        offs.size, gfb.ctypes.data_as(u64p), ng, out.ctypes.data_as(u8p), status.ctypes.data_as(u8p)))
    return out, status[:offs.size]


def sketch_genomes(genomes, k, p, canon=True, device=0) -> np.ndarray:
    return sketch_batch(*records_layout(genomes), k, p, canon, device)


def hostpack(ascii_bases) -> tuple:
    """ASCII bases -> (codes uint32[ceil(n/16)], valid uint16[ceil(n/16)]) with the library's host-side packer (no device needed)."""
    a = _np(ascii_bases, np.uint8)
    ng = (a.size + 15) // 16
    codes, valid = np.zeros(ng, np.uint32), np.zeros(ng, np.uint16)
    lib.db200_hostpack(a.ctypes.data_as(vp), a.size, codes.ctypes.data_as(C.POINTER(C.c_uint32)), valid.ctypes.data_as(C.POINTER(C.c_uint16)))
    return codes, valid


class Sketcher:
    """Streaming S1 form: add_record() per FASTA record, finish() per genome."""

    def __init__(self, p, k, canon=True, device=0, nslots=1):
        self.p, self.h = p, vp()
        _check(lib.db200_sketcher_create(p, k, int(canon), device, nslots, C.byref(self.h)))

    def add_record(self, slot, rec: bytes):
        _check(lib.db200_sketcher_add_record(self.h, slot, rec, len(rec)))

    def finish(self, slot) -> np.ndarray:
        out = np.zeros(1 << self.p, dtype=np.uint8)
        _check(lib.db200_sketcher_finish(self.h, slot, out.ctypes.data_as(u8p)))
        return out

    def close(self):
        if self.h and lib is not None:        # (module globals are gone at interpreter shutdown)
            lib.db200_sketcher_destroy(self.h)
            self.h = vp()

    __del__ = close


class PackedGenomes:
    """Device-resident 2-bit genome store (the sketch kernel's HBM input format)."""

    def __init__(self, bases, rec_offsets, genome_rec_begin, k, device=0):
        """bases: uint8 numpy array (host) or an int device pointer to ASCII bases."""
        if isinstance(bases, int):
            bases_ptr = vp(bases)
        else:
            bases = _np(bases, np.uint8)
            bases_ptr = bases.ctypes.data_as(vp)
        offs = _np(rec_offsets, np.uint64)
        grb = _np(genome_rec_begin, np.uint64)
        self.h = vp()
        self.ngenomes = grb.size - 1
        _check(lib.db200_pack_genomes(device, bases_ptr, offs.ctypes.data_as(u64p), offs.size - 1,
                                      grb.ctypes.data_as(u64p), self.ngenomes, k, C.byref(self.h)))
        self._stats()

    def repack(self, bases, rec_offsets, genome_rec_begin, k):
        """Another batch into the same store (db200_repack_genomes): device buffers are reused."""
        if isinstance(bases, int):
            bases_ptr = vp(bases)
        else:
            bases = _np(bases, np.uint8)
            bases_ptr = bases.ctypes.data_as(vp)
        offs = _np(rec_offsets, np.uint64)
        grb = _np(genome_rec_begin, np.uint64)
        self.ngenomes = grb.size - 1
        _check(lib.db200_repack_genomes(self.h, bases_ptr, offs.ctypes.data_as(u64p), offs.size - 1, grb.ctypes.data_as(u64p), self.ngenomes, k))
        self._stats()

    def _stats(self):
        pb, km, nb = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(lib.db200_packed_genomes_stats(self.h, C.byref(pb), C.byref(km), C.byref(nb)))
        self.packed_bytes, self.kmers, self.nbases = pb.value, km.value, nb.value

    def sketch_dev(self, p, canon, d_registers: int, stream: int = 0):
        _check(lib.db200_sketch_packed_dev(self.h, p, int(canon), vp(d_registers), vp(stream)))

    def close(self):
        if self.h and lib is not None:        # (module globals are gone at interpreter shutdown)
            lib.db200_packed_genomes_free(self.h)
            self.h = vp()

    __del__ = close


# ---- cardinalities / all-pairs ----------------------------------------------------------------
def cardinalities(regs, p, estim=ERTL_MLE, device=0) -> np.ndarray:
    regs = _np(regs, np.uint8).reshape(-1, 1 << p)
    out = np.zeros(regs.shape[0], dtype=np.float64)
    _check(lib.db200_cardinalities(device, regs.ctypes.data_as(u8p), regs.shape[0], p, estim, out.ctypes.data_as(f64p)))
    return out


def union(regs, p, device=0) -> np.ndarray:
    """Element-wise maximum of the rows (hll_t::operator+= folded, src/union.cpp:33-58)."""
    regs = _np(regs, np.uint8).reshape(-1, 1 << p)
    out = np.zeros(1 << p, dtype=np.uint8)
    _check(lib.db200_union(device, regs.ctypes.data_as(u8p), regs.shape[0], p, out.ctypes.data_as(u8p)))
    return out


def compress(regs, p, new_p, device=0) -> np.ndarray:
    """hll_t::compress(new_p) (hll.h:903-924) of every row -> uint8[n][2^new_p]."""
    regs = _np(regs, np.uint8).reshape(-1, 1 << p)
    out = np.zeros((regs.shape[0], 1 << max(min(new_p, p), 0)), dtype=np.uint8)
    _check(lib.db200_compress(device, regs.ctypes.data_as(u8p), regs.shape[0], p, new_p, out.ctypes.data_as(u8p)))
    return out


def dist_symmetric(regs, p, k=31, estim=ERTL_MLE, jestim=ERTL_MLE, result_type=JI, order=ORDER_ROW_FIRST, device=0,
                   row_begin=0, row_end=None, out=None, card=None) -> np.ndarray:
    regs = _np(regs, np.uint8).reshape(-1, 1 << p)
    n = regs.shape[0]
    re_ = n if row_end is None else min(row_end, n)
    tri = lambda r: (r * (2 * n - r - 1)) // 2
    npairs = tri(re_) - tri(row_begin)
    if out is None:
        out = np.zeros(npairs, dtype=np.float32)
    if card is not None and np.size(card) != n:
        raise ValueError(f"card holds {np.size(card)} values for {n} sketches")
    prm = dist_params(p, k, estim, jestim, result_type, order, card=card)
    _check(lib.db200_dist_symmetric_rows(device, regs.ctypes.data_as(u8p), n, C.byref(prm), row_begin, re_, out.ctypes.data_as(f32p)))
    return out


def dist_symmetric_stream(regs, p, on_rows, k=31, estim=ERTL_MLE, jestim=ERTL_MLE, result_type=JI, order=ORDER_ROW_FIRST, device=0,
                          row_begin=0, row_end=None, block_pairs=0, card=None):
    """Row-block streaming: on_rows(row_begin, row_end, values float32[]) per block (values are copied for the caller);
    a non-zero / raising callback aborts the call."""
    regs = _np(regs, np.uint8).reshape(-1, 1 << p)
    n = regs.shape[0]
    err = []

    def tramp(_ud, rb, re_, vals, nv):
        try:
            r = on_rows(int(rb), int(re_), np.ctypeslib.as_array(vals, shape=(int(nv),)).copy() if nv else np.zeros(0, np.float32))
            return int(r or 0)
        except BaseException as e:   # noqa: BLE001 - must not propagate through the C frame
            err.append(e)
            return 1
    cb = ROWS_CB(tramp)
    prm = dist_params(p, k, estim, jestim, result_type, order, card=card)
    rc = lib.db200_dist_symmetric_stream(device, regs.ctypes.data_as(u8p), n, C.byref(prm), row_begin, n if row_end is None else row_end,
                                         block_pairs, cb, None)
    if err:
        raise err[0]
    _check(rc)


def dist_rect(refs, qrys, p, k=31, estim=ERTL_MLE, jestim=ERTL_MLE, result_type=JI, device=0, card=None, card_queries=None) -> np.ndarray:
    refs = _np(refs, np.uint8).reshape(-1, 1 << p)
    qrys = _np(qrys, np.uint8).reshape(-1, 1 << p)
    out = np.zeros((qrys.shape[0], refs.shape[0]), dtype=np.float32)
    prm = dist_params(p, k, estim, jestim, result_type, ORDER_COL_FIRST, card=card, card_queries=card_queries)
    _check(lib.db200_dist_rect(device, refs.ctypes.data_as(u8p), refs.shape[0], qrys.ctypes.data_as(u8p), qrys.shape[0],
                               C.byref(prm), out.ctypes.data_as(f32p)))
    return out


# db200_neighbor == the reference's validx_t = std::pair<float, uint32_t> (src/sketch_and_cmp.h:605)
NEIGHBOR_DTYPE = np.dtype([("value", np.float32), ("index", np.uint32)])
DIST_MEASURES = (MASH_DIST, FULL_MASH_DIST, CONTAINMENT_DIST, FULL_CONTAINMENT_DIST, SYMMETRIC_CONTAINMENT_DIST)


def knn_symmetric(regs, p, nneighbors, k=31, estim=ERTL_MLE, jestim=ERTL_MLE, result_type=JI, order=ORDER_COL_FIRST,
                  device=0, card=None) -> np.ndarray:
    """-> structured array [n][nneighbors] of (value, index), best first (nndist_loop, src/sketch_and_cmp.h:712-783)."""
    regs = _np(regs, np.uint8).reshape(-1, 1 << p)
    out = np.zeros((regs.shape[0], nneighbors), dtype=NEIGHBOR_DTYPE)
    prm = dist_params(p, k, estim, jestim, result_type, order, card=card)
    _check(lib.db200_dist_knn_symmetric(device, regs.ctypes.data_as(u8p), regs.shape[0], C.byref(prm), nneighbors, vp(out.ctypes.data)))
    return out


def knn_rect(refs, qrys, p, nneighbors, k=31, estim=ERTL_MLE, jestim=ERTL_MLE, result_type=JI, device=0, card=None, card_queries=None) -> np.ndarray:
    refs = _np(refs, np.uint8).reshape(-1, 1 << p)
    qrys = _np(qrys, np.uint8).reshape(-1, 1 << p)
    out = np.zeros((qrys.shape[0], nneighbors), dtype=NEIGHBOR_DTYPE)
    prm = dist_params(p, k, estim, jestim, result_type, ORDER_COL_FIRST, card=card, card_queries=card_queries)
    _check(lib.db200_dist_knn_rect(device, refs.ctypes.data_as(u8p), refs.shape[0], qrys.ctypes.data_as(u8p), qrys.shape[0],
                                   C.byref(prm), nneighbors, vp(out.ctypes.data)))
    return out


class DistPlan:
    """Device-resident all-pairs plan (threshold bit-planes + cardinalities) over a device register matrix."""

    def __init__(self, device=0):
        self.h = vp()
        _check(lib.db200_dist_plan_create(device, C.byref(self.h)))

    def prepare_dev(self, d_regs: int, n: int, p: int, estim=ERTL_MLE, stream: int = 0):
        _check(lib.db200_dist_plan_prepare_dev(self.h, vp(d_regs), n, p, estim, vp(stream)))

    def begin_dev(self, n: int, p: int, estim, reg_min: int, reg_max: int, stream: int = 0):
        _check(lib.db200_dist_plan_begin_dev(self.h, n, p, estim, reg_min, reg_max, vp(stream)))

    def add_rows_dev(self, d_regs: int, row_begin: int, nrows: int, stream: int = 0):
        _check(lib.db200_dist_plan_add_rows_dev(self.h, vp(d_regs), row_begin, nrows, vp(stream)))

    def finish_dev(self):
        _check(lib.db200_dist_plan_finish_dev(self.h))

    def run_symmetric_dev(self, prm: DistParams, row_begin: int, row_end: int, d_out: int, stream: int = 0):
        _check(lib.db200_dist_plan_run_symmetric_dev(self.h, C.byref(prm), row_begin, row_end, vp(d_out), vp(stream)))

    def run_rect_dev(self, prm: DistParams, nr: int, nq: int, d_out: int, stream: int = 0):
        _check(lib.db200_dist_plan_run_rect_dev(self.h, C.byref(prm), nr, nq, vp(d_out), vp(stream)))

    def run_knn_dev(self, prm: DistParams, nr: int, nq: int, nneighbors: int, d_out: int, stream: int = 0):
        _check(lib.db200_dist_plan_run_knn_dev(self.h, C.byref(prm), nr, nq, nneighbors, vp(d_out), vp(stream)))

    def run_knn_rows_dev(self, prm: DistParams, row_begin: int, row_end: int, nneighbors: int, d_out: int, stream: int = 0):
        _check(lib.db200_dist_plan_run_knn_rows_dev(self.h, C.byref(prm), row_begin, row_end, nneighbors, vp(d_out), vp(stream)))

    def cardinalities_dev(self) -> int:
        ptr = vp()
        _check(lib.db200_dist_plan_cardinalities_dev(self.h, C.byref(ptr)))
        return ptr.value

    def last_run_info(self):
        pairs, tiles, thr = C.c_uint64(), C.c_uint64(), C.c_int()
        _check(lib.db200_dist_plan_last_run_info(self.h, C.byref(pairs), C.byref(tiles), C.byref(thr)))
        return pairs.value, tiles.value, thr.value

    def close(self):
        if self.h and lib is not None:        # (module globals are gone at interpreter shutdown)
            lib.db200_dist_plan_destroy(self.h)
            self.h = vp()

    __del__ = close
