/* hinge_b200 — C ABI of the B200-native HINGE hot path.
 *
 * HINGE has no in-process plugin interface: its hot path sits behind three
 * executables (`hinge filter | maximal | layout`, /root/reference/src/hinge:9-17)
 * that talk through files.  This header is the FFI surface a maintainer binds
 * instead of (or from inside) those executables; every entry point names the
 * reference code it replaces.  Plain pointers and sizes only; all integers are
 * int32 unless stated; offsets are int64.  Functions return 0 (HG_OK) or a
 * negative hg_status; nothing throws, nothing falls back to the CPU: without a
 * CUDA device every compute call returns HG_ERR_CUDA.
 *
 * Threading: a context is bound to one device and one stream; use one context
 * per thread / per GPU.  Host buffers passed in are only read during the call.
 */
#ifndef HINGE_B200_H
#define HINGE_B200_H
#include <stdint.h>

#include "../hinge_b200/csrc/hg_params.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hg_ctx hg_ctx;

typedef enum hg_status {
    HG_OK = 0,
    HG_ERR_CUDA = -1,      /* no device / CUDA runtime error (see hg_last_error) */
    HG_ERR_ARG = -2,       /* bad argument or call order */
    HG_ERR_INPUT = -3,     /* overlap records violate the .las invariants */
    HG_ERR_NOMEM = -4,
    HG_ERR_IO = -5,        /* file-level helpers: unreadable DB / .las / INI */
    HG_ERR_NO_ALIGNMENTS = -6 /* reference: "No alignments!", exit 1 (filter.cpp:505-508) */
} hg_status;

enum { HG_MEM_HOST = 0, HG_MEM_DEVICE = 1 };

/* positive, non-fatal: the variable-size annotation pool was too small.  hg_filter_phase3 has
 * already grown it: rerun phases 1-3 (hg_filter does this itself; sharded callers rerun on ALL
 * ranks, see hinge_b200/sharding.py) */
#define HG_RETRY_POOL 1

enum hg_option {
    HG_OPT_KEEP_COVERAGE = 1, /* keep the 40-bp coverage profiles (.coverage.txt) */
    HG_OPT_PROFILE = 2,       /* record CUDA events between the kernels of a stage */
    HG_OPT_SCATTER_SPREAD = 3,/* tuning aid: record windows per warp in the profile scatter (1, 4, 8, 16) */
    HG_OPT_KEEP_MASKS = 5,    /* multi-part runs (--mlas): hg_set_overlaps + hg_filter on the next part keep the
                                 masks the earlier parts computed (reads of later parts still have (0,0)),
                                 as the reference's part loop does (filter.cpp:534,884-889) */
    HG_OPT_ANNO_POOL = 6,     /* initial capacity of the annotation pool (entries; default 2 per owned read + 64 K);
                                 a pool that turns out too small is grown and the stage rerun (HG_RETRY_POOL) */
    HG_OPT_PROFILE_KERNEL = 4 /* tuning aid, form of the coverage-profile kernel: 0 = picked by cut_off and data
                                 shape (20-bp start/end histogram for the nominal cut_off 300 when records
                                 outnumber coverage bins, else the four-event 40-bp form), 1 = always the
                                 four-event form, 3 = the persistent TMA-staged form (bulk copies of the abpos /
                                 aepos columns into shared memory, batches bounded by record volume),
                                 5 / 6 = always the 20-bp form, compiled for 4 / 6 resident CTAs per SM */
};

enum hg_buffer { /* per-read device arrays a sharded run exchanges between phases */
    HG_BUF_MEAN_COV = 1,  /* int32[n_read], -1 = not part of the estimate      */
    HG_BUF_MASK = 2,      /* int32[n_read][2], the .mas intervals               */
    HG_BUF_READ_FLAGS = 3,/* uint8[n_read]                                       */
    HG_BUF_MASK_PACKED = 5, /* uint32[n_read]: both mask bounds in units of gcd(40, tspace), 16 bits each; when
                              bound, phase 2 fills it next to the masks and phase 3 looks B-reads up in it, so a
                              sharded run all-gathers 4 B per read instead of HG_BUF_MASK's 8 (refused with
                              HG_ERR_ARG when a read is too long for 16 bits: bind HG_BUF_MASK then) */
    HG_BUF_MEDIAN_HIST = 4 /* uint32[4098]: histogram of the per-read mean coverage; when bound,
                              phase 1 adds the rank's own reads and the caller sums it across
                              ranks (all-reduce) instead of all-gathering HG_BUF_MEAN_COV */
};

/* ---- context ---------------------------------------------------------- */

/* `stream` is a cudaStream_t (NULL = the legacy default stream). */
int hg_ctx_create(int device, void* stream, hg_ctx** out);
void hg_ctx_destroy(hg_ctx* ctx);
const char* hg_last_error(const hg_ctx* ctx);
const char* hg_version(void);
int hg_set_option(hg_ctx* ctx, int option, int64_t value);
/* Device address and size of a per-read array (rows of reads outside the
 * context's [a_lo, a_hi) are the caller's to fill, e.g. by an NCCL all-gather). */
int hg_device_buffer(hg_ctx* ctx, int which, void** dptr, int64_t* bytes);
/* Make the context use caller-owned device memory for a per-read array (so a
 * framework tensor can be all-gathered in place); call after hg_set_reads. */
int hg_bind_buffer(hg_ctx* ctx, int which, void* dptr, int64_t bytes);
/* Device time of the filter kernels of the last run (needs HG_OPT_PROFILE):
 * ms[0] coverage estimate, [1] median, [2] mask + annotation, [3] hinge calls. */
int hg_filter_kernel_times(hg_ctx* ctx, float* ms, int n);
/* Number of kernels this library has launched so far (process-wide). */
int64_t hg_launch_count(void);

/* ---- inputs ----------------------------------------------------------- */

/* Reads of the (trimmed) DAZZ_DB: lengths and the optional `qual` track.
 * Replaces LAInterface::openDB/getRead/getQV as used by the stages
 * (lib/LAInterface.cpp:133-185,1195-1286,4369-4494); only rlen and the QV tiles
 * are consumed.  qv_off/qv may be NULL (no track => coverage mask only,
 * filter.cpp:304-305).  Host pointers. */
int hg_set_reads(hg_ctx* ctx, int32_t n_read, const int32_t* rlen, const int64_t* qv_off,
                 const uint8_t* qv, int32_t tspace);

/* Overlap records as a struct of arrays, sorted by (aread, bread, abpos) like
 * a LAsort-ed .las.  Replaces LAInterface::getOverlap + LOverlap
 * (lib/LAInterface.cpp:1519-1634): B coordinates are passed exactly as stored
 * in the file (complement-strand coordinates when flags&1); the flip to the
 * forward strand happens on the device.  `diffs` is part of the layout but no
 * stage reads it (may be NULL).  trace_off/trace (raw (diff,bdelta) bytes, tbytes
 * = 1 or 2 per value) are needed by hg_maximal/hg_layout only and may be NULL
 * for hg_filter.  `where` = HG_MEM_HOST copies to the device on the context's
 * stream; HG_MEM_DEVICE adopts 16-byte aligned device pointers without a copy
 * (they must stay valid until replaced).
 *
 * A context may own just a slice of the reads: pass the records whose aread is
 * in [a_lo, a_hi) and that range; n_read arrays stay global. */
int hg_set_overlaps(hg_ctx* ctx, int64_t novl, const int32_t* aread, const int32_t* bread,
                    const int32_t* abpos, const int32_t* aepos, const int32_t* bbpos,
                    const int32_t* bepos, const int32_t* diffs, const int32_t* flags,
                    const int64_t* trace_off, const uint8_t* trace, int32_t tbytes, int32_t where,
                    int32_t a_lo, int32_t a_hi);

/* A context that owns a slice of the reads sees only its own records; the reference's
 * r_begin / r_end are the first and last A-read of the WHOLE .las (filter.cpp:516-517) and
 * decide which record-less reads still get a mask / a coverage-estimate entry.  Call after
 * hg_set_overlaps with the global values (min / max over the shards); it stays in force for
 * later hg_set_overlaps calls whose records lie inside it, until the next hg_set_reads. */
int hg_set_global_range(hg_ctx* ctx, int32_t first_aread, int32_t last_aread);

/* ---- phase exchange of sharded contexts through NVLink peer memory ------ */

/* Instead of phase-level calls with NCCL collectives in between, the contexts of all ranks can
 * be connected once; hg_filter on each of them (called collectively) then runs as ONE stream of
 * kernels: the histogram kernel stores its part into every rank's exchange block, K2 stores
 * every packed mask word into every rank's array, and arrival flags in peer memory replace the
 * barriers (a rank that never arrives turns into HG_ERR_CUDA after 4 s, not a hung GPU).
 *   hg_peer_export   allocates this rank's exchange block; one-process-per-GPU callers get a
 *                    CUDA IPC handle (HG_PEER_HANDLE_BYTES) to pass to the other ranks
 *   hg_peer_connect  maps the blocks of all ranks: `handles` = world x HG_PEER_HANDLE_BYTES in
 *                    rank order (the own entry is ignored); synchronise the ranks afterwards,
 *                    before the first hg_filter
 *   hg_peer_connect_local  the same for `world` contexts of ONE process (peer access, no IPC)
 * Needs 16-bit packable masks (see HG_BUF_MASK_PACKED) and world <= 16. */
#define HG_PEER_HANDLE_BYTES 64
int hg_peer_export(hg_ctx* ctx, int32_t rank, int32_t world, void* handle_out);
int hg_peer_connect(hg_ctx* ctx, const void* handles);
int hg_peer_connect_local(hg_ctx** ctxs, int32_t world);
/* The masks of ALL reads as this rank holds them after a peer-connected run (2 ints per read). */
int hg_peer_masks(hg_ctx* ctx, int32_t* mask);

/* ---- hinge filter (filter.cpp:529-1098) -------------------------------- */

typedef struct hg_filter_summary {
    int32_t r_begin, r_end;  /* first / last A-read with records (filter.cpp:516-517) */
    int32_t cov_est;         /* median of per-read mean coverage (filter.cpp:660-671) */
    int32_t min_cov;         /* max(min_cov, cov_est/3)      (filter.cpp:677-678) */
    int64_t n_annotations;   /* "Number of hinges before filtering" */
    int64_t n_hinges;        /* "Number of hinges" (reads r_begin..r_end-1) */
    float ms_device;         /* device time of the whole stage, CUDA events */
    int32_t n_exact_order;   /* annotations that needed the order-exact sort path */
} hg_filter_summary;

/* One call = the whole stage on the device, no host round trip in between:
 * coverage profiles + estimate -> masks -> repeat annotation -> hinge calls. */
int hg_filter(hg_ctx* ctx, const hg_filter_params* params, hg_filter_summary* out);

/* The same stage split at its two global dependencies, for contexts that own
 * a slice of the reads: exchange the named device array between the calls
 * (all-reduce HG_BUF_MEDIAN_HIST -- or all-gather HG_BUF_MEAN_COV --, then
 * all-gather HG_BUF_MASK_PACKED -- or HG_BUF_MASK).  hg_filter == phase1; phase2;
 * phase3 on one context. */
int hg_filter_phase1(hg_ctx* ctx, const hg_filter_params* params); /* -> HG_BUF_MEDIAN_HIST / HG_BUF_MEAN_COV */
int hg_filter_phase2(hg_ctx* ctx);                                 /* -> HG_BUF_MASK_PACKED / HG_BUF_MASK   */
int hg_filter_phase3(hg_ctx* ctx, hg_filter_summary* out);         /* hinge calls        */

/* hg_filter without the wait at its end, for callers that have host work to overlap or run the stage
 * several times on resident records: hg_filter_enqueue only launches -- phase 1, 2 and 3
 * of a context that owns all reads or exchanges through peer memory (hg_peer_connect) --, hg_filter_finish
 * waits for the stage enqueued last and returns its summary (HG_RETRY_POOL as from hg_filter_phase3).
 * hg_filter == enqueue; finish (+ the rerun after HG_RETRY_POOL). */
int hg_filter_enqueue(hg_ctx* ctx, const hg_filter_params* params);
int hg_filter_finish(hg_ctx* ctx, hg_filter_summary* out);

/* Results of the last filter run, copied to caller-owned host arrays (any may
 * be NULL).  mask/cmask: 2 ints per read (.mas / .cmas lines, filter.cpp:775-788);
 * flags: bit0 = .cov.flag, bit1 = .self.flag; anno_off: n_read+1 offsets into
 * anno_pos/anno_type (.repeat.txt); hinge_keep[k] = 1 if annotation k is a
 * called hinge (.hinges.txt). */
int hg_filter_fetch(hg_ctx* ctx, int32_t* mask, int32_t* cmask, uint8_t* flags, int64_t* anno_off,
                    int32_t* anno_pos, int32_t* anno_type, uint8_t* hinge_keep);
/* Coverage profiles at 40 bp (the .coverage.txt payload, filter.cpp:599-602);
 * needs HG_OPT_KEEP_COVERAGE set before hg_filter.  cov_off gets n_read+1
 * offsets, cov one int per bin; call with cov == NULL first to learn *n_bins. */
int hg_filter_coverage(hg_ctx* ctx, int64_t* cov_off, int32_t* cov, int64_t* n_bins);

/* ---- maximal reads (maximal.cpp:524-878) ------------------------------- */

/* mask: 2 ints per read (the .mas content); NULL = use the masks of the last
 * hg_filter on this context.  maximal_out: n_read bytes, 1 = read survives
 * (.max); contained_by (may be NULL): containing read per read or -1
 * (.contained.txt). */
int hg_maximal(hg_ctx* ctx, const hg_layout_params* params, const int32_t* mask,
               uint8_t* maximal_out, int32_t* contained_by, float* ms_device);

/* The same stage for contexts that own a slice of the reads.  Containment is a recurrence over
 * ascending read ids (maximal.cpp:809,853): a read with an active container of higher id is
 * removed, one without containers survives, the others depend on the final state of containers
 * with LOWER ids, which may live in other shards.  Phase 1 settles what is local and leaves, on
 * the device, the per-read states (uint8[n_read]: own reads 0 unknown / 1 survives / 2 removed,
 * other reads 0) and the lists of the unknown reads (unk: int32[4] per read = read, first list
 * entry, entries, 0; pool: int32 ids of lower-id containers); counts[0..1] = entries used.  The
 * caller MAX-all-reduces the states and all-gathers unk / pool with fixed strides (the
 * "all-gather of the maximal-read bitmap" of BASELINE.json's north_star, over NCCL); phase 2 runs
 * the resolve on every rank and returns the bitmap of ALL reads.  All pointers but counts /
 * counts_all / maximal_out are device pointers. */
int hg_maximal_phase1(hg_ctx* ctx, const hg_layout_params* params, const int32_t* mask, void* state_out,
                      void* unk_out, int64_t unk_cap, void* pool_out, int64_t pool_cap, int32_t* counts);
int hg_maximal_phase2(hg_ctx* ctx, void* state_all, const void* unk_all, const int32_t* counts_all,
                      int32_t world, int64_t unk_stride, const void* pool_all, int64_t pool_stride,
                      uint8_t* maximal_out);

/* ---- layout: candidate extensions + best-overlap selection ------------- */

typedef struct hg_edge {      /* one line of .edges.hinges (hinging.cpp:188-248) */
    int32_t a, b, length, comp, type, weight;
    int32_t eff_a[2], eff_b[2];   /* trimmed match on A / B            */
    int32_t read_a[2], read_b[2]; /* effective (masked) read intervals */
    int32_t raw_a[2], raw_b[2];   /* untrimmed match, B on its forward strand */
    int32_t hinge_pos;            /* .edges.hinges2 */
} hg_edge;

/* Everything `hinge layout` computes from the records: classification of the
 * top-two overlaps per maximal pair (hinging.cpp:473-602), weight ordering
 * (:1066-1071), hinge bookkeeping and hinge graph (:1180-1691) and the
 * best-overlap scoring loop (:1911-2148).  Inputs are the inter-stage files'
 * content: mask (2 ints/read), maximal (1 byte/read), repeat annotations and
 * hinges as CSR (off has n_read+1 entries; pos/type per entry).  Results are
 * fetched with hg_layout_edges (the file-level driver prints every candidate list).
 * ms_device: device time of the kernels alone (CUDA events around every stretch of launches; the
 * host steps between them -- allocation, list-size round trips, union-find -- are not in it). */
int hg_layout(hg_ctx* ctx, const hg_layout_params* params, const int32_t* mask,
              const uint8_t* maximal, const int64_t* rep_off, const int32_t* rep_pos,
              const int32_t* rep_type, const int64_t* hin_off, const int32_t* hin_pos,
              const int32_t* hin_type, float* ms_device);
int hg_layout_edges(hg_ctx* ctx, hg_edge* edges, int64_t capacity, int64_t* n_edges);

/* The same stage for contexts that own a slice of the reads: every rank classifies the pairs of its
 * own reads and picks their edges; mask, maximal and the annotation / hinge lists are the global
 * arrays (gathered before: the filter's results of all shards, the bitmap of hg_maximal_phase2).
 * What crosses shards in between is small and travels as HOST arrays through the caller's collectives:
 *   phase 1 -> contained_out[n_read]: reads found contained after all (hinging.cpp:598-601)  MAX all-reduce
 *   phase 2 -> alive_out[n_hinges]:   0 = hinge killed by a match of an own read (:1262-1321) MIN all-reduce
 *              *n_graph records of the hinge graph (hg_layout_graph, :1365-1640)            all-gather
 *   phase 3    components of the WHOLE hinge graph (:1644-1691), best extension of the own reads;
 *              hg_layout_edges then returns this rank's edges (rank order = read order) */
typedef struct hg_graph_rec { int32_t owner, seq, f[4], flag, rev, u, v; } hg_graph_rec;
int hg_layout_phase1(hg_ctx* ctx, const hg_layout_params* params, const int32_t* mask, const uint8_t* maximal,
                     const int64_t* rep_off, const int32_t* rep_pos, const int32_t* rep_type,
                     const int64_t* hin_off, const int32_t* hin_pos, const int32_t* hin_type,
                     uint8_t* contained_out);
int hg_layout_phase2(hg_ctx* ctx, const uint8_t* contained_all, uint8_t* alive_out, int64_t* n_graph);
int hg_layout_graph(hg_ctx* ctx, hg_graph_rec* out, int64_t capacity);
int hg_layout_phase3(hg_ctx* ctx, const uint8_t* alive_all, const hg_graph_rec* graph_all, int64_t n_graph_all,
                     float* ms_device);

/* ---- file-level drivers: what the three executables do ------------------ */

/* Same flags, inputs, outputs and exit conventions as Reads_filter,
 * get_maximal_reads and hinging (filter.cpp:168, maximal.cpp:238,
 * hinging.cpp:616).  Return value is the process exit code. */
int hg_main_filter(int argc, char** argv);
int hg_main_maximal(int argc, char** argv);
int hg_main_layout(int argc, char** argv);
/* For a process that exits right after one hg_main_* call (the `hinge` executable): leave the CUDA context
 * and its allocations to the process exit instead of tearing them down one by one (the results are on the
 * host by then; the driver reclaims everything).  Off by default: a library user keeps clean teardown. */
void hg_main_exit_after(int on);

#ifdef __cplusplus
}
#endif
#endif
