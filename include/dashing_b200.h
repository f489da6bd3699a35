/* dashing_b200.h — C ABI of the B200-native HyperLogLog engine (libdashing_b200.so).
 *
 * The reference (dnbaker/dashing) has no plugin / FFI interface: its two hot paths are C++
 * templates called in-process.  This header therefore declares exactly the seams a maintainer
 * would cut at the reference's own call sites (paths relative to the reference tree):
 *
 *   S1  sketch      body of the OpenMP genome loop in sketch_core      src/sketch_and_cmp.h:496-520
 *                   and in dist_sketch_and_cmp phase A                 src/sketch_and_cmp.h:338-352
 *                   (today: Encoder::for_each(lambda addh, path, kseq)  bonsai/include/bonsai/encoder.h:509-529,
 *                    hll_t::addh                                         bonsai/hll/include/sketch/hll.h:843-846)
 *   S2  sizes       the serial cardinality loop                        src/sketch_and_cmp.h:377-383
 *                   (hll_t::report -> sum_counts + calculate_estimate   hll.h:773-803, :515-532, :199-246)
 *   S3  all pairs   dist_loop / perform_core_op / partdist_loop        src/sketch_and_cmp.h:785-880, :699-710,
 *                   / dm::parallel_fill oracle                          src/dashing.h:660-712, distmat/distmat.h:459-512
 *                   (result_cmp src/dashing.h:568-592 -> hll_t::jaccard_index / full_set_comparison
 *                    hll.h:1174-1183, :1165-1173, ertl_joint :636-684, ertl_ml_estimate :567-627)
 *
 * Conventions
 *   - Plain C: pointers + sizes, no C++/torch types.  Every function returns 0 on success and a
 *     non-zero DB200_E* code on failure; db200_last_error() returns the calling thread's message.
 *     (The reference's convention is print-and-exit(1), bonsai/include/bonsai/util.h:547-554; the
 *     host wrapper turns a non-zero return into UNRECOVERABLE_ERROR(db200_last_error()).)
 *   - There is NO CPU fallback inside the library: without a usable CUDA device every compute
 *     entry point fails with DB200_ENODEV.  Flag combinations the GPU path does not cover are
 *     rejected with DB200_EUNSUPPORTED so the host keeps using the reference's own code for them.
 *   - Register matrices are row-major uint8_t[n][2^p] (what hll_t::data() yields, hll.h:1029).
 *   - Host-pointer entry points own all device memory and streams internally and are blocking.
 *     `_dev` entry points take DEVICE pointers plus a cudaStream_t (passed as void*), enqueue
 *     work on that stream and return without synchronising, so callers can bracket them with
 *     CUDA events (bench.py) or chain them after a collective (multi-GPU driver).
 *   - CUDA device: an entry point that takes a `device` (or an object bound to one) makes that device current on the calling
 *     thread and leaves it current, as CUDA libraries do; db200_host_alloc / db200_host_free / db200_hostpack / db200_last_error /
 *     db200_kernel_launches never change it.  A host that drives several GPUs from one thread re-selects its device after a call.
 */
#ifndef DASHING_B200_H
#define DASHING_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DB200_VERSION 101
#if defined(__GNUC__)
#define DB200_API __attribute__((visibility("default")))
#else
#define DB200_API
#endif

enum db200_status {
    DB200_OK = 0,
    DB200_EINVAL = 1,       /* bad argument */
    DB200_EUNSUPPORTED = 2, /* valid for the reference, not covered by the GPU path */
    DB200_ENODEV = 3,       /* no CUDA device / driver */
    DB200_ECUDA = 4,        /* CUDA runtime error (message has the details) */
    DB200_ENOMEM = 5
};

/* sketch::hll::EstimationMethod — hll.h:61-65 */
enum db200_estim { DB200_ORIGINAL = 0, DB200_ERTL_IMPROVED = 1, DB200_ERTL_MLE = 2 };
/* sketch::hll::JointEstimationMethod — hll.h:76-81; values 0..2 mean "union + estim", 3 = Ertl joint MLE */
enum db200_jestim { DB200_ERTL_JOINT_MLE = 3 };
/* bns::EmissionType — src/enums.h:13-23 */
enum db200_emission {
    DB200_MASH_DIST = 0, DB200_JI = 1, DB200_SIZES = 2, DB200_FULL_MASH_DIST = 3,
    DB200_FULL_CONTAINMENT_DIST = 4, DB200_CONTAINMENT_INDEX = 5, DB200_CONTAINMENT_DIST = 6,
    DB200_SYMMETRIC_CONTAINMENT_INDEX = 7, DB200_SYMMETRIC_CONTAINMENT_DIST = 8,
    /* extension (not an EmissionType): the union cardinality itself — hll_t::union_size (hll.h:1125-1141), what the
     * reference's Python helper union_size_matrix returns (bonsai/hll/python/util.cpp:164) */
    DB200_UNION_SIZE = 9
};
/* Operand order of result_cmp in the symmetric loop: the reference's TSV/PHYLIP path evaluates
 * cmp(sketches[j], sketches[i]) for j > i (perform_core_op, src/sketch_and_cmp.h:702) while its
 * binary paths evaluate cmp(sketches[i], sketches[j]) (:829, :849).  Only the joint MLE is
 * sensitive to it (last ulp). */
enum db200_order { DB200_ORDER_ROW_FIRST = 0 /* cmp(s[i], s[j]) */, DB200_ORDER_COL_FIRST = 1 /* cmp(s[j], s[i]) */ };

typedef struct db200_dist_params {
    int32_t p;           /* log2 registers per sketch; GPU path: 7 <= p <= 20 */
    int32_t k;           /* k-mer length (only enters Mash/containment distances through ksinv = (float)(1./k)) */
    int32_t estim;       /* enum db200_estim: estimator behind creport() and the union estimate */
    int32_t jestim;      /* DB200_ERTL_JOINT_MLE or anything else for the union path */
    int32_t result_type; /* enum db200_emission */
    int32_t order;       /* enum db200_order (symmetric mode only) */
    /* Cached cardinalities (optional, NULL = evaluate them from the registers under `estim`).  The reference's hll_t carries a
     * cached estimate (value_): a sketch loaded from a file keeps what hll_t::read() computed under the FILE's estimation
     * method (csum(), hll.h:1078) — or the value stored in the file — and the pair loop uses that for the per-sketch terms
     * (creport(), hll.h:780-783) even when the command line names another estimator.  A host that loaded such sketches
     * passes their cached values here (host pointers, read during the call): `card` holds one double per sketch of the
     * symmetric / knn_symmetric matrix, or per REFERENCE of the rectangular forms; `card_queries` one per query.
     * Ignored by the `_dev` plan entry points (a plan's cardinalities are its own). */
    const double *card;
    const double *card_queries;
} db200_dist_params;

/* S4 multi-GPU: every HOST-pointer compute entry point accepts device = DB200_ALL_DEVICES and then
 * shards its units over all visible GPUs, one host thread per device, no change to callers: genomes (sketch_batch,
 * sketcher slots), sketches (cardinalities), block rows of the packed triangle balanced by pair count
 * (dist_symmetric[_rows]) and queries (dist_rect, dist_knn_rect).  The host already holds every input, so each device
 * receives what it needs by H2D and writes its own slice of the output; there is no device-to-device exchange in this
 * form (the multi-process form — one rank per GPU, NCCL all-gather of register shards — is dashing_b200/multigpu.py
 * over the `_dev` entry points).  dist_knn_symmetric runs on device 0 (its tie rule is order dependent).
 * Device-resident objects (packed stores, plans) belong to one device and take a real index. */
#define DB200_ALL_DEVICES (-1)

/* ------------------------------------------------------------------------------------------ */
DB200_API const char *db200_last_error(void);
DB200_API int db200_version(void);
/* Number of usable CUDA devices (0 when there is no driver/GPU; never fails).  DB200_VIRTUAL_DEVICES=N in the
 * environment exposes N logical devices mapped round-robin onto the physical ones (testing the multi-device paths on
 * a single-GPU box). */
DB200_API int db200_device_count(void);
/* Optional: creates the CUDA context of `device` now (seconds on a multi-GPU node) instead of inside the first compute
 * call, so a host can overlap it with file parsing from another thread. */
DB200_API int db200_warmup(int device);
/* Page-locked host memory, so host-pointer entry points can DMA straight from caller buffers. */
DB200_API int db200_host_alloc(void **out, size_t bytes);
DB200_API int db200_host_free(void *ptr);

/* ---- S1: sketching ------------------------------------------------------------------------
 * Streaming form, a drop-in for `enc.for_each([&](u64 kmer){h.addh(kmer);}, path, kseq)`:
 * the host keeps gz/FASTA parsing (kseq) and hands every record over exactly as kseq_read yields
 * it (ks->seq.s, ks->seq.l); records are never joined (encoder.h:444).  `slot` selects one of
 * `nslots` independent sketches (one per OpenMP worker in the reference's loop); calls on
 * distinct slots may come from different threads concurrently.
 * finish() writes the slot's 2^p registers to host memory — the caller copies them into
 * hll_t::mutable_core() (hll.h:1028) and calls not_ready() (:1012) — and clears the slot.
 * GPU path: 1 <= k <= 32, 7 <= p <= 24, unspaced, unwindowed, DNA4 alphabet, WangHash. */
typedef struct db200_sketcher db200_sketcher;
DB200_API int db200_sketcher_create(int p, int k, int canon, int device, uint32_t nslots, db200_sketcher **out);
DB200_API int db200_sketcher_add_record(db200_sketcher *h, uint32_t slot, const char *bases, uint64_t len);
DB200_API int db200_sketcher_finish(db200_sketcher *h, uint32_t slot, uint8_t *registers_out);
DB200_API int db200_sketcher_destroy(db200_sketcher *h);

/* Batch form (what the OpenMP genome loop of sketch_core becomes): genome g owns records
 * [genome_rec_begin[g], genome_rec_begin[g+1]); record r is bases[rec_offsets[r] .. rec_offsets[r+1]).
 * registers_out is uint8_t[ngenomes][2^p] in host memory. */
DB200_API int db200_sketch_batch(int device, int p, int k, int canon,
                       const char *bases, const uint64_t *rec_offsets, uint64_t nrecords,
                       const uint64_t *genome_rec_begin, uint64_t ngenomes, uint8_t *registers_out);

/* Device-side FASTA parsing — the step BEFORE S1 moved onto the GPU as well.  Instead of kseq records the
 * host hands over RAW file text (what read() or gz inflate produced: header lines, newlines, CR and all); the library
 * applies kseq_read's record rules per byte on the device (bonsai/klib/kseq.h:177-218; dashing_b200/csrc/fasta.cuh), packs
 * and sketches.  File f is text[file_off[f] .. file_off[f] + file_len[f]); file_off must be multiples of
 * DB200_FASTA_ALIGN, ascending, at least one block apart and non-overlapping (what lies between files is ignored); genome g
 * is the files [genome_file_begin[g], genome_file_begin[g+1]) folded into one sketch (FNAME_SEP paths, src/substrs.h:7-26).
 * file_status_out (optional, one byte per file): 0 = parsed; 1 = the file shows FASTQ record syntax ('@' headers, '+'
 * lines) which this path does not cover — the registers of ITS GENOME are then unspecified and the host must sketch that
 * genome through the record interface (db200_sketch_batch) instead. */
#define DB200_FASTA_ALIGN 8192
DB200_API int db200_sketch_fasta_batch(int device, int p, int k, int canon, const char *text, const uint64_t *file_off,
                                       const uint64_t *file_len, uint64_t nfiles, const uint64_t *genome_file_begin, uint64_t ngenomes,
                                       uint8_t *registers_out, uint8_t *file_status_out);

/* The host-side half of the batch upload (dashing_b200/csrc/hostpack.cpp): ASCII bases -> the packed store's 2-bit code words
 * (16 bases per uint32, base j at bits [2j, 2j+1], A0 C1 G2 T3) and validity bits (16 bases per uint16; ACGTacgt valid —
 * alph::DNA4, bonsai/include/bonsai/alphabet.h:128).  db200_sketch_batch runs it on a pool of host threads for a share of
 * every batch so that the link carries 0.375 bytes per base for that share (DB200_HOST_PACK=0 disables it; pure CPU code,
 * usable and testable without a device).  codes / valid hold ceil(nbases / 16) entries; a partial last group is zero padded.
 * db200_hostpack_isa(): "avx512bw" | "avx2" | "scalar" — the variant picked at run time. */
DB200_API void db200_hostpack(const uint8_t *ascii, size_t nbases, uint32_t *codes, uint16_t *valid);
DB200_API const char *db200_hostpack_isa(void);

/* Device-resident form used by bench.py (`value`) and the multi-GPU driver.  The packed genome
 * store is the HBM-resident input format of the sketch kernel: 2-bit bases (64 per 16-byte word),
 * a validity bit-plane and a record-start bit-plane (DESIGN.md "Data layout"). */
typedef struct db200_packed_genomes db200_packed_genomes;
/* Packs ASCII (same arguments as db200_sketch_batch) into a device-resident store.  `bases` is a host
 * pointer (pinned memory is DMA'd directly) or, under unified addressing, a device pointer. */
DB200_API int db200_pack_genomes(int device, const char *bases, const uint64_t *rec_offsets, uint64_t nrecords,
                       const uint64_t *genome_rec_begin, uint64_t ngenomes, int k, db200_packed_genomes **out);
/* Packs another batch into an EXISTING store (its device buffers are reused and only grow): what a host that streams batch
 * after batch through one GPU calls instead of free + pack (cudaMalloc / cudaFree of GB-sized buffers cost tens of ms).
 * Any sketch of the store's previous content must have completed. */
DB200_API int db200_repack_genomes(db200_packed_genomes *g, const char *bases, const uint64_t *rec_offsets, uint64_t nrecords,
                         const uint64_t *genome_rec_begin, uint64_t ngenomes, int k);
DB200_API int db200_packed_genomes_free(db200_packed_genomes *g);
/* Bytes the sketch kernel streams for this store (2-bit + validity + start planes) and its k-mer count. */
DB200_API int db200_packed_genomes_stats(const db200_packed_genomes *g, uint64_t *packed_bytes, uint64_t *kmers, uint64_t *bases);
/* Sketches every genome of the store into d_registers (device, uint8_t[ngenomes][2^p], overwritten). */
DB200_API int db200_sketch_packed_dev(const db200_packed_genomes *g, int p, int canon, uint8_t *d_registers, void *stream);

/* ---- S2: per-sketch cardinalities ---------------------------------------------------------- */
DB200_API int db200_cardinalities(int device, const uint8_t *regs, uint64_t n, int p, int estim, double *out);

/* ---- set operations: `dashing union` and `dashing fold` ---------------------------------------
 * db200_union: element-wise maximum of n register arrays — hll_t::operator+= (hll.h:958-992) folded over the inputs
 * as union_core does (src/union.cpp:33-58).  out: uint8_t[2^p].  n == 0 gives the empty sketch.
 * db200_compress: hll_t::compress(new_p) (hll.h:903-924) applied to each of n sketches; out: uint8_t[n][2^new_p].
 * new_p == p copies, new_p > p is the reference's "Can't compress to a larger size" error (DB200_EINVAL). */
DB200_API int db200_union(int device, const uint8_t *regs, uint64_t n, int p, uint8_t *out);
DB200_API int db200_compress(int device, const uint8_t *regs, uint64_t n, int p, int new_p, uint8_t *out);

/* ---- S3: all-pairs ------------------------------------------------------------------------
 * Symmetric mode: out is the packed upper triangle in distmat order,
 *   idx(i,j) = i(2n-i-1)/2 + j-i-1 for i < j   (distmat/distmat.h:260-276),
 * n(n-1)/2 floats; db200_dist_symmetric_rows computes rows [row_begin,row_end) only and writes
 * them contiguously starting at out[0] (rows are contiguous in distmat order), which is how the
 * multi-GPU driver shards the triangle by block-row.
 * Rectangular mode (-Q/-F, partdist_loop): out[q*nr + j] = result_cmp(refs[j], queries[q]). */
DB200_API int db200_dist_symmetric(int device, const uint8_t *regs, uint64_t n, const db200_dist_params *prm, float *out);
DB200_API int db200_dist_symmetric_rows(int device, const uint8_t *regs, uint64_t n, const db200_dist_params *prm,
                              uint64_t row_begin, uint64_t row_end, float *out);
DB200_API int db200_dist_rect(int device, const uint8_t *ref_regs, uint64_t nr, const uint8_t *qry_regs, uint64_t nq,
                    const db200_dist_params *prm, float *out);

/* Row-block streaming form ("optional row-block callback lets the host keep its async TSV / PHYLIP / binary
 * writers unchanged").  Rows [row_begin, row_end) of the symmetric matrix are computed in blocks of whole rows holding about
 * `block_pairs` values (0 = 8 Mi) and handed to `cb` in ascending row order: values[0 .. nvalues) are rows [rb, re) in distmat
 * order (row i contributes its n-1-i pairs), valid only during the call.  The kernel and the device->host copy of the next
 * blocks run while the callback works, and no buffer of n(n-1)/2 floats exists on either side (100,000 sketches: 20 GB).
 * A non-zero return from the callback aborts the call.  With DB200_ALL_DEVICES every device streams a contiguous row range
 * from its own host thread: the callback is then entered concurrently (in row order per device, not globally) and must be
 * thread-safe — e.g. pwrite() into the distmat file at offset 9 + 4 * idx(rb, rb + 1). */
typedef int (*db200_rows_cb)(void *user, uint64_t row_begin, uint64_t row_end, const float *values, uint64_t nvalues);
DB200_API int db200_dist_symmetric_stream(int device, const uint8_t *regs, uint64_t n, const db200_dist_params *prm,
                                          uint64_t row_begin, uint64_t row_end, uint64_t block_pairs, db200_rows_cb cb, void *user);

/* Device-resident form.  A plan owns the derived HBM structures (threshold bit-planes,
 * per-sketch cardinalities and value ranges); prepare() builds them from a device register
 * matrix, run_*() enqueue the all-pairs kernel.  `stream` is a cudaStream_t. */
typedef struct db200_dist_plan db200_dist_plan;
DB200_API int db200_dist_plan_create(int device, db200_dist_plan **out);
DB200_API int db200_dist_plan_destroy(db200_dist_plan *pl);
/* d_regs: device uint8_t[n][2^p].  Synchronises once (reads back the global register range). */
DB200_API int db200_dist_plan_prepare_dev(db200_dist_plan *pl, const uint8_t *d_regs, uint64_t n, int p, int estim, void *stream);
/* The same in three steps, for a driver that overlaps the plane build with the exchange of the register shards (one rank per
 * GPU: dashing_b200/multigpu.py::allgather_prepare_overlapped).  begin: allocations for n sketches whose registers all lie in
 * [reg_min, reg_max] (the exact global range, e.g. an all-reduce of the ranks' local minima / maxima — a looser range only
 * costs thresholds).  add_rows: threshold planes, counts, tails and cardinalities of rows [row_begin, row_begin + nrows) of the
 * matrix at d_regs (the base of all n rows), enqueued on `stream` — call it for each shard as soon as it has landed.
 * finish: after every row has been added (no synchronisation). */
DB200_API int db200_dist_plan_begin_dev(db200_dist_plan *pl, uint64_t n, int p, int estim, int reg_min, int reg_max, void *stream);
DB200_API int db200_dist_plan_add_rows_dev(db200_dist_plan *pl, const uint8_t *d_regs, uint64_t row_begin, uint64_t nrows, void *stream);
DB200_API int db200_dist_plan_finish_dev(db200_dist_plan *pl);
/* Rows [row_begin,row_end) of the symmetric matrix into d_out (device floats, rows contiguous from d_out[0]). */
DB200_API int db200_dist_plan_run_symmetric_dev(db200_dist_plan *pl, const db200_dist_params *prm,
                                      uint64_t row_begin, uint64_t row_end, float *d_out, void *stream);
/* Rectangular: sketches [0,nr) of the prepared matrix are references, [nr, nr+nq) queries. */
DB200_API int db200_dist_plan_run_rect_dev(db200_dist_plan *pl, const db200_dist_params *prm, uint64_t nr, uint64_t nq,
                                 float *d_out, void *stream);
/* ---- S3b: k nearest neighbours (--nearest-neighbors; nndist_loop / perform_nns, src/sketch_and_cmp.h:642-783) ----
 * Per sketch, the `nneighbors` best (value, index) pairs over all OTHER sketches (symmetric) or, per query, over all
 * references (rect), best first: ascending for the distance measures (MASH_DIST, FULL_MASH_DIST, *CONTAINMENT_DIST),
 * descending for the similarity measures (emt2nntype, src/dashing.h:268-280).  The all-pairs values stay in HBM; only
 * n x nneighbors pairs come back.  db200_neighbor is layout-compatible with the reference's validx_t =
 * std::pair<float, uint32_t> (:605).  Semantics are those of the reference run with one thread (and of its -Q/-F mode
 * with any number of threads): sketches are visited in ascending index and a value replaces the current worst only if
 * STRICTLY better, which decides which of several equal values at the cut survive; unused slots keep the reference's
 * (+-FLT_MAX, UINT32_MAX) filler.  Values are result_cmp(..., ksinv = 1./k in double) as at :729 (the all-pairs paths
 * round ksinv to float first), evaluated as cmp(sketches[j], sketches[i]) for j > i when prm->order is
 * DB200_ORDER_COL_FIRST (what :670 does).  1 <= nneighbors <= 1024. */
typedef struct db200_neighbor { float value; uint32_t index; } db200_neighbor;
DB200_API int db200_dist_knn_symmetric(int device, const uint8_t *regs, uint64_t n, const db200_dist_params *prm, uint32_t nneighbors,
                             db200_neighbor *out /* [n][nneighbors] */);
DB200_API int db200_dist_knn_rect(int device, const uint8_t *ref_regs, uint64_t nr, const uint8_t *qry_regs, uint64_t nq,
                        const db200_dist_params *prm, uint32_t nneighbors, db200_neighbor *out /* [nq][nneighbors] */);
/* Device form on a prepared plan: nq == 0 -> symmetric over the plan's n sketches (nr ignored), else references
 * [0,nr) and queries [nr, nr+nq) as in db200_dist_plan_run_rect_dev.  d_out: device db200_neighbor[rows][nneighbors]. */
DB200_API int db200_dist_plan_run_knn_dev(db200_dist_plan *pl, const db200_dist_params *prm, uint64_t nr, uint64_t nq, uint32_t nneighbors,
                                db200_neighbor *d_out, void *stream);

/* Partial symmetric table: only the pairs (i, j > i) with i in [row_begin, row_end) are visited (ascending i, as always).  For the
 * DISTANCE measures the retained set is the nneighbors smallest (value, index) keys, so the tables of disjoint row ranges merge
 * exactly (per row: the nneighbors smallest keys of their union) — what dashing_b200/multigpu.py does across ranks.  For the
 * similarity measures the reference's retained set depends on the visiting order and partial tables do not merge. */
DB200_API int db200_dist_plan_run_knn_rows_dev(db200_dist_plan *pl, const db200_dist_params *prm, uint64_t row_begin, uint64_t row_end,
                                               uint32_t nneighbors, db200_neighbor *d_out, void *stream);

/* Device pointer to the plan's per-sketch cardinalities (double[n]) — valid until the next prepare/destroy. */
DB200_API int db200_dist_plan_cardinalities_dev(db200_dist_plan *pl, const double **d_card);
/* Launch accounting for bench.py: kernels launched by this library since process start. */
DB200_API uint64_t db200_kernel_launches(void);
/* Name and algorithmic-byte model of the last all-pairs run (for the roofline object). */
DB200_API int db200_dist_plan_last_run_info(const db200_dist_plan *pl, uint64_t *pairs, uint64_t *tiles, int *thresholds);

#ifdef __cplusplus
}
#endif
#endif /* DASHING_B200_H */
