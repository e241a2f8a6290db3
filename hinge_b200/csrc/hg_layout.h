// Host-visible declarations of the maximal / layout kernels' launchers.
#ifndef HG_LAYOUT_H
#define HG_LAYOUT_H
#include <cuda_runtime.h>
#include <stdint.h>

#include "hg_params.h"

namespace hg {

struct RecView;
struct ReadView;

struct KeyIdx2 {
    int key, idx;
};
struct KeyIdx2Greater {
    __host__ __device__ bool operator()(const KeyIdx2& a, const KeyIdx2& b) const { return a.key > b.key; }
};

// One classified overlap between two maximal reads that extends A
// (FORWARD / FORWARD_INTERNAL / BACKWARD / BACKWARD_INTERNAL).
struct Cand {
    int a, b, type, comp, weight, length;
    int eas, eae, ebs, ebe;  // trimmed match
    int as, ae, bs, be;      // raw match, B on its forward strand
    int64_t rec;             // record index (trace, file order)
    int rank, pad;           // 0 / 1: best or second-best of its pair
};

struct PairOut {
    int* counters;          // [0] big pairs [1] sort scratch used [3] overflow [4] pairs [5] cands
    int64_t* big_pairs;     // first record of pairs with > 16 records
    int big_cap;
    KeyIdx2* sort_scratch;
    int sort_cap;
    int4* pairs;            // layout: (a, b, first record lo31, hi)
    int pair_cap;
    Cand* cands;
    int cand_cap;
    uint8_t* contained_flag;  // layout: per read, "[contained] Should not happen"
};

// containment on compact lists (k_contain_lists / k_contain_resolve)
struct ContainIO {
    const uint8_t* active0;  // n_read: long enough to take part (maximal.cpp:541-547)
    const uint8_t* rtype;    // per record: class of the top-two records of every pair, 255 otherwise
    uint8_t* state;          // n_read: 0 unknown, 1 survives (maximal), 2 removed
    int4* unk;               // unknown reads: (read, first list entry, list entries, -)
    int unk_cap;
    int* pool;               // lists of lower-id containing reads
    int pool_cap;
    int* counters;           // [0] unknown reads [1] pool used [2] overflow
};

void launch_classify_reads(const RecView& rv, const ReadView& rd, const hg_layout_params& P, const int2* mask,
                           const uint8_t* active, int sort_passes, uint8_t* rtype, const PairOut& po,
                           cudaStream_t st);
void launch_contain_lists(const RecView& rv, const ReadView& rd, const ContainIO& io, cudaStream_t st);
void launch_contain_resolve(const int4* unk, const int* counts, int world, int unk_stride, const int* pool,
                            int pool_stride, uint8_t* state, int* sweeps_out, cudaStream_t st);

void launch_classify(const RecView& rv, const ReadView& rd, const hg_layout_params& P,
                     const int2* mask, const uint8_t* active, int mode, int sort_passes,
                     uint8_t* rtype, const PairOut& po, cudaStream_t st);
void launch_contain_init(const RecView& rv, const ReadView& rd, const uint8_t* active0,
                         const uint8_t* rtype, uint8_t* state, cudaStream_t st);
void launch_contain_step(const RecView& rv, const ReadView& rd, const uint8_t* active0,
                         const uint8_t* rtype, uint8_t* state, int* remaining, cudaStream_t st);

// ---- layout selection -------------------------------------------------------

struct HingeView {           // CSR over reads, device
    const int64_t* off;      // n_read + 1
    const int* pos;
    const int* type;
};

struct GraphRec {            // one .hgraph line (+ the union it implies)
    int owner, seq;          // emitting read and order within it
    int f[4];                // the four leading integers of the line
    int flag, rev;           // 1 = hinge-hinge edge, 0 = hinge-killed hinge
    int u, v;                // hinge node ids (flag == 1)
};
struct NkRec {               // new_killed_hinges_vec[owner].push_back(...)
    int owner, seq, pos, type;
};
struct SkipRec {             // one .edges.skipped line
    int owner, seq, cand;
};

struct SelectIO {
    int n_read;
    const uint8_t* active;   // reads
    const Cand* cands;
    const int4* ranges;      // per read: fwd [x,y), bwd [z,w) into order[]
    int4* ranges_out;        // written by k_order_candidates
    int* order;              // candidate indices, sorted by weight per list
    KeyIdx2* sort_scratch;   // one per candidate
    HingeView hv, kv, nk;    // hinges, killed hinges, new killed hinges
    uint8_t* hinge_alive;    // per hinge entry: kill pass result, later AND component size
    int* counters;           // [0] graph recs [1] nk recs [2] skip recs [3] overflow
    GraphRec* graph;
    int graph_cap;
    NkRec* nkout;
    int nk_cap;
    SkipRec* skips;
    int skip_cap;
    int2* chosen;            // per read x 2 (fwd, bwd): (candidate index or -1, hinge_pos)
};

// device-resident pair / candidate lists of the layout stage (k_layout_pairs, k_order_candidates)
struct LayoutLists {
    int2* pair_ref;          // per read: (first pair slot, pairs with an active B)
    int2* cand_ref;          // per read: (first candidate slot, candidates)
    int* bkt_ref;            // per read: first slot of its hash-bucket scratch
    int2* pairs;             // per pair: (B, first candidate relative to the read's block << 2 | candidates)
    Cand* cands;
    int* hash_next;          // scratch of hash_iteration_order: one int per pair ...
    int* hash_out;
    int* hash_bkt;           // ... and one per bucket
    uint8_t* contained_flag; // per read: "[contained] Should not happen" (hinging.cpp:598)
    int* counters;           // [0] pair slots [1] candidate slots [2] bucket slots [3] overflow [4] sort scratch
    KeyIdx2* sort_scratch;   // pairs with more than 16 records
    int sort_cap;
    const int* grow_at;      // bucket-count schedule of std::unordered_map (hash_growth_schedule)
    const int* grow_bkt;
    int ngrow;
};

void launch_layout_count_pairs(const RecView& rv, const ReadView& rd, const uint8_t* active, int2* pair_ref,
                               unsigned long long* total, cudaStream_t st);
void launch_layout_pairs(const RecView& rv, const ReadView& rd, const hg_layout_params& P, const int2* mask,
                         const uint8_t* active, const LayoutLists& L, cudaStream_t st);
void launch_order_candidates(const LayoutLists& L, const SelectIO& io, cudaStream_t st);
void launch_apply_contained(int n, const uint8_t* contained, uint8_t* active, int* count, cudaStream_t st);
void launch_gather_chosen(int n_read, const int2* chosen, const Cand* cands, Cand* out, int2* out_ref, int* count,
                          cudaStream_t st);
void launch_sort_candidates(const SelectIO& io, cudaStream_t st);
void launch_hinge_graph(const RecView& rv, const hg_layout_params& P, const SelectIO& io,
                        cudaStream_t st);
void launch_best_extension(const hg_layout_params& P, const SelectIO& io, cudaStream_t st);

}  // namespace hg
#endif
