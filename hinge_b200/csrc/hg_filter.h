// Host-visible declarations of the filter-stage kernels' launchers.
#ifndef HG_FILTER_H
#define HG_FILTER_H
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "hg_params.h"

namespace hg {

struct RecView;
struct ReadView;
struct MaskAnnoOut;
struct PeerView;

enum : uint8_t { kFlagCov = 1, kFlagSelf = 2, kFlagSkipHinge = 4 };

constexpr int kFlatBins = 4096;     // histogram words (= coverage bins) per CTA of the flat kernels
constexpr int kFlatMaxReads = 512;  // reads per batch

// Device buffers of one filter run (all sized at hg_set_reads / hg_set_overlaps).
struct FilterScratch {
    int num_sms = 148;
    // ingest
    int* self_cnt = nullptr;                // n_read: records with A == B
    // K1 (profile build)
    int* cov_maxbin = nullptr;              // n_read: last bin of the cut-off-free profile
    int* counters1 = nullptr;               // 8: [0] reads handed to the per-read fallback
    int* mean_cov = nullptr;                // n_read, -1 = not part of the estimate
    unsigned int* med_hist = nullptr;       // 4097 + 4096
    int* scal = nullptr;                    // 8: cov_est, MIN_COV, radix state
    // K2
    int2* mask = nullptr;      // n_read
    uint32_t* mask_pk = nullptr;  // n_read, caller-owned (HG_BUF_MASK_PACKED): both bounds in units of mask_g
    int mask_g = 1;
    int2* cmask = nullptr;     // n_read
    uint8_t* rflags = nullptr; // n_read
    int2* anno_ref = nullptr;  // n_read
    int2* anno_pool = nullptr;
    int anno_cap = 0;
    int* counters = nullptr;   // 8
    int4* work_items = nullptr;  // 3 x n_read: K4 work items (push_work_item)
    int* big_list = nullptr;   // n_read
    int* exact_list = nullptr; // n_read: reads whose hinge calls need the exact sort order
    int2* flat_batch = nullptr;       // (first read, histogram words in use) per batch (flat_nbatch + 1)
    int* flat_rbase = nullptr;        // per read, first histogram word inside its batch (-1: fallback path)
    uint32_t* flat_prof = nullptr;    // scanned packed profiles, kFlatBins words per batch (K1 -> K2)
    int* flat_rbatch = nullptr;       // per read: its batch (-1 = outside the planned range)
    int4* flat_desc = nullptr;        // per batch, two entries: (first read, reads, words, 0) (first record lo, hi, records, 0)
    uint16_t* flat_zmap = nullptr;    // K2: bit maps of the batches' bins (kFlatBins bits per batch each)
    uint16_t* flat_cmap = nullptr;
    int flat_lo = 0, flat_hi = 0;     // planned read range
    int flat_nbatch = 0;
    int flat_spread = 8;              // record windows per warp in the scatter (tuning aid)
    int flat_kernel = 0;              // form of K1.  0: first / second by cut-off and shape, 1: first, 3: third
                                      // (k_profile_tma; first when the columns are not 16-byte aligned),
                                      // 5 / 6: second form compiled for 4 / 6 resident CTAs per SM
    bool flat_capped = false;         // the plan bounds the batches by record volume (third form)
    double flat_bins_total = 0;       // coverage bins of the planned reads
    float flat_bins_per_record = 0;   // ... per record of the context
    unsigned long long* big_scratch = nullptr;
    int big_slot_words = 0, big_warps = 0;
    // K4
    uint8_t* hinge_keep = nullptr;  // anno_cap
    uint8_t* hinge_scratch = nullptr;
    int hinge_cap = 0, hinge_warps = 0;
    int4* item_log = nullptr;       // HG_OPT_PROFILE: (read, cycles, support, exact n) per work item
    cudaStream_t side_stream[3] = {nullptr, nullptr, nullptr};  // the size tiers of the exact-order kernels run side by side
    cudaEvent_t side_event[4] = {nullptr, nullptr, nullptr, nullptr};
};

void launch_csr_validate(const RecView& rv, const ReadView& rd, int64_t* read_off, int* self_cnt, int* err,
                         cudaStream_t st);
void launch_qv_mask(int n_read, const int64_t* qv_off, const uint8_t* qv, int tspace, int2* out,
                    cudaStream_t st);
void launch_median(const ReadView& rd, const hg_filter_params& P, FilterScratch& s, int mode,
                   cudaStream_t st);
void launch_mask_anno(const RecView& rv, const ReadView& rd, const hg_filter_params& P,
                      int r_begin, int r_end, FilterScratch& s, int* cov0, const int64_t* cov0_off,
                      const PeerView& peer, cudaStream_t st);
// flat kernels (hg_filter_flat.cu): host-side batch plan + launchers of the two phases
struct FlatPlan {
    std::vector<int2> batch;   // (first read, histogram words in use) per batch, closed by (hi, 0)
    std::vector<int> rbase;    // per read: first word of its profile inside its batch, -1 = fallback path
    std::vector<int> rbatch;   // per read: its batch, -1 = outside [lo, hi)
    std::vector<int4> desc;    // per batch: (first read, reads, words, 0) (first record lo, hi, records, 0)
};
void flat_plan(const int* rlen, const int64_t* read_off, int lo, int hi, int n_read, int cut_off, bool cap_records,
               FlatPlan* plan);
void launch_profile(const RecView& rv, const ReadView& rd, const hg_filter_params& P, int r_begin,
                    int r_end, FilterScratch& s, cudaStream_t st);
void launch_mask_anno_flat(const RecView& rv, const ReadView& rd, const hg_filter_params& P, int r_begin,
                           int r_end, FilterScratch& s, const MaskAnnoOut& out, cudaStream_t st);
void launch_hinge_call(const RecView& rv, const ReadView& rd, const hg_filter_params& P,
                       FilterScratch& s, const PeerView& peer, cudaStream_t st);
// peer exchange (sharded runs, hg_peer_connect): push the rank's histogram part / pick the median
// from all parts / publish the masks K2 stored into the other ranks' arrays
void launch_peer_hist_push(FilterScratch& s, const PeerView& peer, cudaStream_t st);
void launch_peer_median_pick(const hg_filter_params& P, FilterScratch& s, const PeerView& peer, cudaStream_t st);
void launch_peer_signal_masks(FilterScratch& s, const PeerView& peer, cudaStream_t st);
void launch_debug_warp_sort(void* data, const int* off, int count, int descending, int* g, int* l, void* tmp,
                            cudaStream_t st);
void launch_max_pileup(const int64_t* read_off, int n_read, int* out_max, cudaStream_t st);

}  // namespace hg
#endif
