// The context behind the C ABI: device buffers, stream, per-stage scratch.
#ifndef HG_CTX_H
#define HG_CTX_H
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/hinge_b200.h"
#include "hg_device.cuh"
#include "hg_filter.h"

namespace hg {
struct LayoutResult;
void free_layout_result(LayoutResult* r);
struct LayoutRun;
void free_layout_run(LayoutRun* r);
}

struct hg_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int num_sms = 148;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    // reads
    int n_read = 0, tspace = 100, max_rlen = 0, rlen_q999 = 0;
    bool has_qv = false;
    int* d_rlen = nullptr;
    int2* d_qvmask = nullptr;
    std::vector<int> h_rlen;

    // overlaps (owned copies unless adopted)
    int64_t novl = 0;
    bool adopted = false;
    int32_t *d_aread = nullptr, *d_bread = nullptr, *d_abpos = nullptr, *d_aepos = nullptr;
    int32_t *d_bbpos = nullptr, *d_bepos = nullptr, *d_flags = nullptr;
    int64_t* d_trace_off = nullptr;
    uint8_t* d_trace = nullptr;
    int tbytes = 1;
    bool has_trace = false;
    int64_t* d_read_off = nullptr;
    int* d_err = nullptr;
    int a_lo = 0, a_hi = 0, r_begin = 0, r_end = -1, max_pileup = 0;
    int64_t cap_novl = 0, cap_trace = 0;  // capacity of the owned record buffers
    size_t cap_hinge_scratch = 0;

    // filter
    hg::FilterScratch fs;
    hg_filter_params fp;
    bool filter_params_set = false, filter_done = false;
    int shape_version = 0, configured_shape = -1;
    int reads_version = 0, plan_reads_version = -1, plan_cut_off = 0, plan_lo = 0, plan_hi = 0, plan_shape = -1;  // flat plan
    int keep_cov = 0;
    bool keep_masks = false;  // HG_OPT_KEEP_MASKS
    int anno_pool_hint = 0;   // HG_OPT_ANNO_POOL
    int* d_cov0 = nullptr;
    int64_t* d_cov0_off = nullptr;
    std::vector<int64_t> h_cov0_off;

    // per-kernel timing (HG_OPT_PROFILE): events between the launches of a stage
    static constexpr int kMarks = 16;
    bool profile = false;
    cudaEvent_t marks[kMarks] = {};
    void mark(int i) {
        if (profile) cudaEventRecord(marks[i], stream);
    }
    bool ext_mean_cov = false, ext_mask = false, ext_med_hist = false;  // bound to caller-owned memory

    // sharded runs: first / last A-read of the whole .las (hg_set_global_range), -1 = unknown
    int g_begin = -1, g_end = -1;
    // phase exchange through peer memory (hg_peer_export / hg_peer_connect)
    hg::PeerView peer{};          // world <= 1: off
    uint8_t* peer_block = nullptr;  // this rank's exchange block
    size_t peer_block_bytes = 0;
    bool peer_ipc[hg::kMaxPeers] = {};  // base[r] came from cudaIpcOpenMemHandle
    bool peer_connected = false;

    // maximal / layout scratch, kept across calls (grown on demand)
    struct StageScratch {
        uint8_t *active0 = nullptr, *state = nullptr, *rtype = nullptr;
        int64_t cap_reads = 0, cap_rtype = 0;
        int4* unk = nullptr;
        int* pool = nullptr;
        int unk_cap = 0, pool_cap = 0;
        int* counters = nullptr;   // 16 ints: [0..7] pair lists, [8..11] containment lists, [12] sweeps
        int64_t* big_pairs = nullptr;
        void* sort_scratch = nullptr;
        int big_cap = 0, sort_cap = 0;
    } ms;

    hg::LayoutResult* layout = nullptr;  // result of the last hg_layout
    hg::LayoutRun* layout_run = nullptr; // between the phases of a sharded hg_layout

    hg::RecView rec_view() const;
    hg::ReadView read_view() const;
};

namespace hg {
extern int64_t g_launches;  // kernels launched by this library (all contexts)
int set_err(hg_ctx* c, int code, const std::string& msg);
int cuda_check(hg_ctx* c, cudaError_t e, const char* what);
template <typename T>
int dev_alloc(hg_ctx* c, T** p, size_t n, const char* what) {
    if (*p) {
        cudaFree(*p);
        *p = nullptr;
    }
    if (n == 0) n = 1;
    return cuda_check(c, cudaMalloc((void**)p, n * sizeof(T)), what);
}
}  // namespace hg

#define HG_TRY(x)                \
    do {                         \
        int _rc = (x);           \
        if (_rc != HG_OK) return _rc; \
    } while (0)

#endif
