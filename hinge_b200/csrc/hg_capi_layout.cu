// C ABI, part 2: hg_maximal, hg_layout, hg_layout_edges (include/hinge_b200.h).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <numeric>
#include <set>
#include <unordered_map>

#include "hg_ctx.h"
#include "hg_layout.h"
#include "hg_layout_result.h"

using namespace hg;

namespace hg {

__global__ void k_length_filter(const int2* __restrict__ mask, int n, int thr,
                                uint8_t* __restrict__ active) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) active[i] = (mask[i].y - mask[i].x < thr) ? 0 : 1;  // maximal.cpp:541-547
}

template <typename T>
struct DevBuf {  // scoped device allocation
    T* p = nullptr;
    size_t n = 0;
    ~DevBuf() { cudaFree(p); }
    int alloc(hg_ctx* c, size_t count, const char* what) {
        cudaFree(p);
        p = nullptr;
        n = count;
        return cuda_check(c, cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(T)), what);
    }
};

struct PairBufs {
    DevBuf<int> counters;
    DevBuf<int64_t> big_pairs;
    DevBuf<KeyIdx2> sort_scratch;
    DevBuf<int4> pairs;
    DevBuf<Cand> cands;
    DevBuf<uint8_t> contained;
    PairOut po;
    int make(hg_ctx* c, int big_cap, int sort_cap, int pair_cap, int cand_cap) {
        HG_TRY(counters.alloc(c, 8, "pair counters"));
        HG_TRY(big_pairs.alloc(c, big_cap, "big pairs"));
        HG_TRY(sort_scratch.alloc(c, sort_cap, "pair sort scratch"));
        HG_TRY(pairs.alloc(c, pair_cap, "pair keys"));
        HG_TRY(cands.alloc(c, cand_cap, "candidates"));
        HG_TRY(contained.alloc(c, c->n_read, "contained flags"));
        po.counters = counters.p; po.big_pairs = big_pairs.p; po.big_cap = big_cap;
        po.sort_scratch = sort_scratch.p; po.sort_cap = sort_cap; po.pairs = pairs.p;
        po.pair_cap = pair_cap; po.cands = cands.p; po.cand_cap = cand_cap;
        po.contained_flag = contained.p;
        return HG_OK;
    }
};

// HINGE_B200_TIMING=1: wall time of the steps inside hg_maximal / hg_layout on stderr (each lap
// synchronises the stream first, so only use it to find out where the time goes)
struct StepTimer {
    bool on = getenv("HINGE_B200_TIMING") != nullptr;
    cudaStream_t st;
    struct timespec t0;
    explicit StepTimer(cudaStream_t s) : st(s) { clock_gettime(CLOCK_MONOTONIC, &t0); }
    void lap(const char* what) {
        if (!on) return;
        cudaStreamSynchronize(st);
        struct timespec t1;
        clock_gettime(CLOCK_MONOTONIC, &t1);
        fprintf(stderr, "[hinge_b200 timing]     %-34s %8.2f ms\n", what,
                1e3 * (double)(t1.tv_sec - t0.tv_sec) + 1e-6 * (double)(t1.tv_nsec - t0.tv_nsec));
        t0 = t1;
    }
};

static int upload_mask(hg_ctx* c, const int32_t* mask) {
    if (mask)
        return cuda_check(c, cudaMemcpyAsync(c->fs.mask, mask, 8ull * c->n_read, cudaMemcpyHostToDevice, c->stream), "mask H2D");
    if (!c->filter_done) return set_err(c, HG_ERR_ARG, "no mask given and no hg_filter run on this context");
    return HG_OK;
}

// Copies the device-resident candidate lists of the last hg_layout to the host (the file writers of
// `hinge layout` print every candidate; the array-level API only needs the chosen ones).
const LayoutResult* layout_result(hg_ctx* c) {
    LayoutResult* R = c->layout;
    if (!R || R->materialized) return R;
    cudaSetDevice(c->device);
    R->cands.resize((size_t)R->n_cand_slots);
    R->order.resize((size_t)R->n_cand_slots);
    R->ranges.resize((size_t)R->n_read);
    cudaMemcpy(R->cands.data(), R->d_cands, sizeof(Cand) * (size_t)R->n_cand_slots, cudaMemcpyDeviceToHost);
    cudaMemcpy(R->order.data(), R->d_order, sizeof(int) * (size_t)R->n_cand_slots, cudaMemcpyDeviceToHost);
    cudaMemcpy(R->ranges.data(), R->d_ranges, sizeof(int4) * (size_t)R->n_read, cudaMemcpyDeviceToHost);
    R->materialized = true;
    return R;
}
void free_layout_result(LayoutResult* r) {
    if (!r) return;
    cudaFree(r->d_cands);
    cudaFree(r->d_order);
    cudaFree(r->d_ranges);
    delete r;
}

}  // namespace hg



// Scratch of the maximal stage, kept in the context: per-read flags, per-record classes, the
// containment lists and the (rarely used) buffers of the order-exact pair sort.
static int maximal_scratch(hg_ctx* c, int unk_cap, int pool_cap, int big_cap, int sort_cap) {
    hg_ctx::StageScratch& m = c->ms;
    const int n = c->n_read;
    if (m.cap_reads < n) {
        HG_TRY(dev_alloc(c, &m.active0, n, "active"));
        HG_TRY(dev_alloc(c, &m.state, n, "state"));
        m.cap_reads = n;
    }
    if (m.cap_rtype < c->novl) {
        HG_TRY(dev_alloc(c, &m.rtype, (size_t)c->novl, "record classes"));
        m.cap_rtype = c->novl;
    }
    if (!m.counters) HG_TRY(dev_alloc(c, &m.counters, 16, "stage counters"));
    if (m.unk_cap < unk_cap) {
        HG_TRY(dev_alloc(c, &m.unk, unk_cap, "unknown reads"));
        m.unk_cap = unk_cap;
    }
    if (m.pool_cap < pool_cap) {
        HG_TRY(dev_alloc(c, &m.pool, pool_cap, "container lists"));
        m.pool_cap = pool_cap;
    }
    if (m.big_cap < big_cap) {
        HG_TRY(dev_alloc(c, &m.big_pairs, big_cap, "big pairs"));
        m.big_cap = big_cap;
    }
    if (m.sort_cap < sort_cap) {
        KeyIdx2* p = (KeyIdx2*)m.sort_scratch;
        HG_TRY(dev_alloc(c, &p, sort_cap, "pair sort scratch"));
        m.sort_scratch = p;
        m.sort_cap = sort_cap;
    }
    return HG_OK;
}

// Classification + containment lists of the context's own reads (everything up to the point where
// a sharded run has to look at other shards).  Leaves state / unk / pool / counters on the device.
// Returns HG_OK, or a positive value when a list overflowed and was grown: call again.
static int maximal_local(hg_ctx* c, const hg_layout_params* P, int* n_unknown, int* pool_used) {
    hg_ctx::StageScratch& m = c->ms;
    cudaStream_t st = c->stream;
    const int n = c->n_read;
    for (int attempt = 0;; attempt++) {
        HG_TRY(maximal_scratch(c, std::max(m.unk_cap, (c->a_hi - c->a_lo) / 4 + 1024),
                               std::max(m.pool_cap, c->a_hi - c->a_lo + 4096), std::max(m.big_cap, 1 << 14),
                               std::max(m.sort_cap, 1 << 18)));
        PairOut po{};
        po.counters = m.counters; po.big_pairs = m.big_pairs; po.big_cap = m.big_cap;
        po.sort_scratch = (KeyIdx2*)m.sort_scratch; po.sort_cap = m.sort_cap;
        ContainIO io{};
        io.active0 = m.active0; io.rtype = m.rtype; io.state = m.state; io.unk = m.unk; io.unk_cap = m.unk_cap;
        io.pool = m.pool; io.pool_cap = m.pool_cap; io.counters = m.counters + 8;
        k_length_filter<<<(n + 255) / 256, 256, 0, st>>>(c->fs.mask, n, P->length_threshold, m.active0);
        cudaMemsetAsync(m.state, 0, n, st);  // reads of other shards: unknown until their states arrive
        // the reference sorts every pair twice before it reads the top two (maximal.cpp:647-654, 791)
        launch_classify_reads(c->rec_view(), c->read_view(), *P, c->fs.mask, m.active0, 2, m.rtype, po, st);
        launch_contain_lists(c->rec_view(), c->read_view(), io, st);
        int cnt[16];
        HG_TRY(cuda_check(c, cudaGetLastError(), "maximal kernels"));
        HG_TRY(cuda_check(c, cudaMemcpyAsync(cnt, m.counters, sizeof cnt, cudaMemcpyDeviceToHost, st), "D2H"));
        HG_TRY(cuda_check(c, cudaStreamSynchronize(st), "maximal: classify + lists"));
        if (!cnt[3] && !cnt[10]) {
            *n_unknown = cnt[8];
            *pool_used = cnt[9];
            return HG_OK;
        }
        if (attempt > 8) return set_err(c, HG_ERR_NOMEM, "maximal: lists kept overflowing");
        // grow what overflowed and run again
        HG_TRY(maximal_scratch(c, std::max(m.unk_cap, cnt[8] + 1024), std::max(m.pool_cap, cnt[9] + 4096),
                               std::max(m.big_cap, cnt[0] + 1024), cnt[3] ? std::max(m.sort_cap * 4, cnt[1] + 1024) : m.sort_cap));
    }
}

extern "C" {

int hg_maximal(hg_ctx* c, const hg_layout_params* P, const int32_t* mask, uint8_t* maximal_out,
               int32_t* contained_by, float* ms_device) {
    if (!c || !P || !maximal_out || c->novl <= 0) return set_err(c, HG_ERR_ARG, "hg_maximal: bad arguments");
    if (!c->has_trace) return set_err(c, HG_ERR_ARG, "hg_maximal needs the trace (pass trace_off/trace to hg_set_overlaps)");
    if (c->a_lo != 0 || c->a_hi != c->n_read)
        return set_err(c, HG_ERR_ARG, "hg_maximal on a context that owns a slice of the reads: containment looks at reads "
                                      "of lower id in other shards; use hg_maximal_phase1 / hg_maximal_phase2");
    cudaSetDevice(c->device);
    cudaStream_t st = c->stream;
    const int n = c->n_read;
    HG_TRY(upload_mask(c, mask));
    StepTimer tm(st);
    cudaEventRecord(c->ev0, st);
    int n_unknown = 0, pool_used = 0;
    HG_TRY(maximal_local(c, P, &n_unknown, &pool_used));
    tm.lap("maximal: classify + containment lists");
    hg_ctx::StageScratch& m = c->ms;
    launch_contain_resolve(m.unk, m.counters + 8, 1, m.unk_cap, m.pool, m.pool_cap, m.state, m.counters + 12, st);
    cudaEventRecord(c->ev1, st);
    std::vector<uint8_t> hstate(n), hact(n);
    HG_TRY(cuda_check(c, cudaGetLastError(), "containment resolve"));
    HG_TRY(cuda_check(c, cudaMemcpyAsync(hstate.data(), m.state, n, cudaMemcpyDeviceToHost, st), "D2H"));
    HG_TRY(cuda_check(c, cudaMemcpyAsync(hact.data(), m.active0, n, cudaMemcpyDeviceToHost, st), "D2H"));
    HG_TRY(cuda_check(c, cudaStreamSynchronize(st), "D2H"));
    tm.lap("maximal: resolve + D2H");
    if (ms_device) cudaEventElapsedTime(ms_device, c->ev0, c->ev1);
    for (int i = 0; i < n; i++) maximal_out[i] = hstate[i] == 1 ? 1 : 0;
    if (contained_by) {
        // .contained.txt names the LAST containing read in std::unordered_map iteration order
        // (maximal.cpp:789-857): replay the key sequence into the real container
        std::vector<uint8_t> ht((size_t)c->novl);
        std::vector<int32_t> hb((size_t)c->novl);
        std::vector<int64_t> hoff((size_t)n + 1);
        HG_TRY(cuda_check(c, cudaMemcpyAsync(ht.data(), m.rtype, (size_t)c->novl, cudaMemcpyDeviceToHost, st), "D2H"));
        HG_TRY(cuda_check(c, cudaMemcpyAsync(hb.data(), c->d_bread, 4ull * c->novl, cudaMemcpyDeviceToHost, st), "D2H"));
        HG_TRY(cuda_check(c, cudaMemcpyAsync(hoff.data(), c->d_read_off, 8ull * (n + 1), cudaMemcpyDeviceToHost, st), "D2H"));
        HG_TRY(cuda_check(c, cudaStreamSynchronize(st), "D2H"));
        for (int i = 0; i < n; i++) {
            contained_by[i] = -1;
            if (!(hact[i] && hstate[i] == 2)) continue;
            std::unordered_map<int, int> um;  // B -> has a BCOVERA record among its top two
            for (int64_t k = hoff[i]; k < hoff[i + 1]; k++) um[hb[k]] = 0;
            for (int64_t k = hoff[i]; k < hoff[i + 1]; k++)
                if (ht[k] == HG_BCOVERA) um[hb[k]] = 1;
            for (auto it = um.begin(); it != um.end(); ++it)
                if (it->second) contained_by[i] = it->first;
        }
        tm.lap("maximal: .contained.txt replay (host)");
    }
    return HG_OK;
}

// ---- sharded form -----------------------------------------------------------------------
// phase 1 (every rank): classification and containment lists of the rank's own reads.
//   state_out   device, uint8[n_read]: own reads filled (0 unknown / 1 survives / 2 removed), the
//               others 0 -- a MAX all-reduce over the ranks assembles the array
//   unk_out     device, int32[4 * unk_cap]; pool_out device, int32[pool_cap]: the lists of this
//               rank's unknown reads, to be all-gathered (fixed strides); counts[0..1] = entries used
// phase 2 (every rank, redundantly): the resolve over the gathered lists; maximal_out (host,
// n_read bytes) is the maximal-read bitmap of ALL reads.
int hg_maximal_phase1(hg_ctx* c, const hg_layout_params* P, const int32_t* mask, void* state_out,
                      void* unk_out, int64_t unk_cap, void* pool_out, int64_t pool_cap, int32_t* counts) {
    if (!c || !P || !state_out || !counts || c->novl <= 0) return set_err(c, HG_ERR_ARG, "hg_maximal_phase1: bad arguments");
    if (!c->has_trace) return set_err(c, HG_ERR_ARG, "hg_maximal needs the trace (pass trace_off/trace to hg_set_overlaps)");
    cudaSetDevice(c->device);
    cudaStream_t st = c->stream;
    HG_TRY(upload_mask(c, mask));
    int n_unknown = 0, pool_used = 0;
    HG_TRY(maximal_local(c, P, &n_unknown, &pool_used));
    counts[0] = n_unknown;
    counts[1] = pool_used;
    hg_ctx::StageScratch& m = c->ms;
    HG_TRY(cuda_check(c, cudaMemcpyAsync(state_out, m.state, c->n_read, cudaMemcpyDeviceToDevice, st), "state"));
    if (unk_out && pool_out && n_unknown <= unk_cap && pool_used <= pool_cap) {
        if (n_unknown) HG_TRY(cuda_check(c, cudaMemcpyAsync(unk_out, m.unk, sizeof(int4) * (size_t)n_unknown, cudaMemcpyDeviceToDevice, st), "lists"));
        if (pool_used) HG_TRY(cuda_check(c, cudaMemcpyAsync(pool_out, m.pool, sizeof(int) * (size_t)pool_used, cudaMemcpyDeviceToDevice, st), "lists"));
    }
    return cuda_check(c, cudaStreamSynchronize(st), "hg_maximal_phase1");
}

int hg_maximal_phase2(hg_ctx* c, void* state_all, const void* unk_all, const int32_t* counts_all, int32_t world,
                      int64_t unk_stride, const void* pool_all, int64_t pool_stride, uint8_t* maximal_out) {
    if (!c || !state_all || !counts_all || world < 1 || !maximal_out)
        return set_err(c, HG_ERR_ARG, "hg_maximal_phase2: bad arguments");
    cudaSetDevice(c->device);
    cudaStream_t st = c->stream;
    int* d_counts = nullptr;
    HG_TRY(dev_alloc(c, &d_counts, 2 * (size_t)world, "counts"));
    int rc = cuda_check(c, cudaMemcpyAsync(d_counts, counts_all, sizeof(int) * 2 * (size_t)world, cudaMemcpyHostToDevice, st), "H2D");
    if (rc == HG_OK) {
        launch_contain_resolve((const int4*)unk_all, d_counts, world, (int)unk_stride, (const int*)pool_all,
                               (int)pool_stride, (uint8_t*)state_all, nullptr, st);
        std::vector<uint8_t> hstate((size_t)c->n_read);
        rc = cuda_check(c, cudaMemcpyAsync(hstate.data(), state_all, c->n_read, cudaMemcpyDeviceToHost, st), "D2H");
        if (rc == HG_OK) rc = cuda_check(c, cudaStreamSynchronize(st), "hg_maximal_phase2");
        if (rc == HG_OK)
            for (int i = 0; i < c->n_read; i++) maximal_out[i] = hstate[i] == 1 ? 1 : 0;
    }
    cudaFree(d_counts);
    return rc;
}

// Bucket-count schedule of this toolchain's std::unordered_map<int, T> (see hash_iteration_order,
// hg_order.h), read off the real container: at[i] = element count whose insertion makes the bucket
// count bkt[i].
static void hash_growth_schedule(int max_elements, std::vector<int>* at, std::vector<int>* bkt) {
    std::unordered_map<int, int> m;
    size_t cur = m.bucket_count();
    at->clear();
    bkt->clear();
    for (int k = 1; k <= max_elements; k++) {
        m[k] = 0;
        if (m.bucket_count() != cur) {
            cur = m.bucket_count();
            at->push_back(k);
            bkt->push_back((int)cur);
        }
    }
}

}  // extern "C"

namespace hg {

// One run of the layout stage: the device buffers that live from the pair pass to the best-extension
// pass.  A context that owns all reads runs the three phases back to back (hg_layout); a sharded run
// exchanges two small per-read / per-hinge arrays and the hinge-graph records between them
// (hg_layout_phase1..3).
struct LayoutRun {
    hg_ctx* c = nullptr;
    hg_layout_params P;
    int n = 0;
    int64_t nh = 0;
    DevBuf<uint8_t> d_active, d_contained, d_alive;
    DevBuf<int2> d_pair_ref, d_cand_ref, d_chosen, d_pairs;
    DevBuf<int> d_bkt_ref, d_cnt, d_grow, d_hpos, d_htype, d_kpos, d_ktype, d_npos, d_ntype, d_hnext, d_hout, d_hbkt;
    DevBuf<int64_t> d_hoff, d_koff, d_noff;
    DevBuf<unsigned long long> d_total;
    DevBuf<KeyIdx2> d_sort, d_bigsort;
    SelectIO io;
    LayoutLists L;
    std::vector<uint8_t> active0;   // before the "[contained]" fix-up
    std::vector<GraphRec> graph;    // this context's records (phase 2)
    std::vector<NkRec> nks;
    int ngrow = 0;
    // device time of the kernels alone: event pairs around every stretch of launches (the host steps
    // between them -- allocation, the list-size round trips, union-find -- are not in it)
    cudaEvent_t seg[8] = {};
    int nseg = 0;
    float kernel_ms = 0;
    void seg_begin() {
        if (nseg + 2 > 8) return;
        for (int i = 0; i < 2; i++)
            if (!seg[nseg + i]) cudaEventCreate(&seg[nseg + i]);
        cudaEventRecord(seg[nseg], c->stream);
    }
    void seg_end() {
        if (nseg + 2 > 8) return;
        cudaEventRecord(seg[nseg + 1], c->stream);
        nseg += 2;
    }
    float kernel_time() {  // call after a synchronisation
        float total = 0;
        for (int i = 0; i + 1 < nseg; i += 2) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, seg[i], seg[i + 1]) == cudaSuccess) total += ms;
        }
        return total;
    }
    ~LayoutRun() {
        for (int i = 0; i < 8; i++)
            if (seg[i]) cudaEventDestroy(seg[i]);
    }

    int h2d(void* d, const void* h, size_t bytes) {
        return bytes ? cuda_check(c, cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, c->stream), "H2D") : HG_OK;
    }
    int phase1(const hg_layout_params* params, const int32_t* mask, const uint8_t* maximal, const int64_t* rep_off,
               const int32_t* rep_pos, const int32_t* rep_type, const int64_t* hin_off, const int32_t* hin_pos,
               const int32_t* hin_type, uint8_t* contained_out);
    int phase2(const uint8_t* contained_all, uint8_t* alive_out);
    int phase3(const uint8_t* alive_all, const GraphRec* graph_all, int64_t n_graph_all);
};

void free_layout_run(LayoutRun* r) { delete r; }

int LayoutRun::phase1(const hg_layout_params* params, const int32_t* mask, const uint8_t* maximal,
                      const int64_t* rep_off, const int32_t* rep_pos, const int32_t* rep_type,
                      const int64_t* hin_off, const int32_t* hin_pos, const int32_t* hin_type,
                      uint8_t* contained_out) {
    P = *params;
    cudaStream_t st = c->stream;
    n = c->n_read;
    free_layout_result(c->layout);
    c->layout = new LayoutResult();
    LayoutResult& R = *c->layout;
    R.n_read = n;
    R.mask.assign(mask, mask + 2 * (size_t)n);
    StepTimer tm(st);

    // ---- who takes part (hinging.cpp:877-913, 954-960, 398-412)
    R.active.assign(n, 1);
    for (int i = 0; i < n; i++) {
        if (P.delete_telomeres && rep_off[i + 1] - rep_off[i] > P.num_events_telomere) R.active[i] = 0;
        if (mask[2 * i + 1] - mask[2 * i] < P.length_threshold) {
            R.active[i] = 0;
            R.garbage.push_back(i);
        }
        R.active[i] = R.active[i] && maximal[i];
    }
    active0 = R.active;
    // ---- hinges, killed hinges = annotations that are not hinges (hinging.cpp:1180-1197)
    R.hin_off.assign(hin_off, hin_off + n + 1);
    R.hin_pos.assign(hin_pos, hin_pos + hin_off[n]);
    R.hin_type.assign(hin_type, hin_type + hin_off[n]);
    R.kil_off.assign((size_t)n + 1, 0);
    R.kil_pos.clear();
    R.kil_type.clear();
    for (int i = 0; i < n; i++) {
        for (int64_t k = rep_off[i]; k < rep_off[i + 1]; k++) {
            bool is_hinge = false;
            for (int64_t h = hin_off[i]; h < hin_off[i + 1] && !is_hinge; h++)
                is_hinge = hin_pos[h] == rep_pos[k] && hin_type[h] == rep_type[k];
            if (!is_hinge) {
                R.kil_pos.push_back(rep_pos[k]);
                R.kil_type.push_back(rep_type[k]);
            }
        }
        R.kil_off[i + 1] = (int64_t)R.kil_pos.size();
    }
    nh = hin_off[n];
    const int64_t nkil = R.kil_off[n];
    std::vector<int> grow_at, grow_bkt;
    hash_growth_schedule(std::max(c->max_pileup, 16) + 2, &grow_at, &grow_bkt);
    ngrow = (int)grow_at.size();
    tm.lap("layout: active flags, killed-hinge lists (host)");

    HG_TRY(upload_mask(c, mask));
    HG_TRY(d_active.alloc(c, n, "active")); HG_TRY(d_contained.alloc(c, n, "contained flags"));
    HG_TRY(d_pair_ref.alloc(c, n, "pair refs")); HG_TRY(d_cand_ref.alloc(c, n, "candidate refs"));
    HG_TRY(d_bkt_ref.alloc(c, n, "bucket refs")); HG_TRY(d_cnt.alloc(c, 16, "counters"));
    HG_TRY(d_total.alloc(c, 1, "pair total")); HG_TRY(d_grow.alloc(c, 2 * grow_at.size() + 2, "hash schedule"));
    HG_TRY(d_hoff.alloc(c, n + 1, "hinge off")); HG_TRY(d_hpos.alloc(c, nh, "hinge pos")); HG_TRY(d_htype.alloc(c, nh, "hinge type"));
    HG_TRY(d_koff.alloc(c, n + 1, "killed off")); HG_TRY(d_kpos.alloc(c, nkil, "killed pos")); HG_TRY(d_ktype.alloc(c, nkil, "killed type"));
    HG_TRY(d_noff.alloc(c, n + 1, "nk off")); HG_TRY(d_alive.alloc(c, nh, "hinge alive"));
    HG_TRY(d_chosen.alloc(c, 2 * (size_t)n, "chosen"));
    HG_TRY(h2d(d_active.p, R.active.data(), n));
    HG_TRY(h2d(d_grow.p, grow_at.data(), 4 * grow_at.size()));
    HG_TRY(h2d(d_grow.p + grow_at.size(), grow_bkt.data(), 4 * grow_bkt.size()));
    HG_TRY(h2d(d_hoff.p, R.hin_off.data(), 8 * ((size_t)n + 1)));
    HG_TRY(h2d(d_hpos.p, R.hin_pos.data(), 4 * (size_t)nh));
    HG_TRY(h2d(d_htype.p, R.hin_type.data(), 4 * (size_t)nh));
    HG_TRY(h2d(d_koff.p, R.kil_off.data(), 8 * ((size_t)n + 1)));
    HG_TRY(h2d(d_kpos.p, R.kil_pos.data(), 4 * (size_t)nkil));
    HG_TRY(h2d(d_ktype.p, R.kil_type.data(), 4 * (size_t)nkil));
    cudaMemsetAsync(d_contained.p, 0, n, st);
    cudaMemsetAsync(d_pair_ref.p, 0, sizeof(int2) * (size_t)n, st);  // reads of other shards: no pairs here
    cudaMemsetAsync(d_cand_ref.p, 0, sizeof(int2) * (size_t)n, st);
    tm.lap("layout: uploads");

    // ---- K5: pairs between maximal reads, their top two classified, candidates in the reference's order
    cudaEventRecord(c->ev0, st);
    seg_begin();
    launch_layout_count_pairs(c->rec_view(), c->read_view(), d_active.p, d_pair_ref.p, d_total.p, st);
    seg_end();
    unsigned long long total_pairs = 0;
    HG_TRY(cuda_check(c, cudaMemcpyAsync(&total_pairs, d_total.p, 8, cudaMemcpyDeviceToHost, st), "D2H"));
    HG_TRY(cuda_check(c, cudaStreamSynchronize(st), "count pairs"));
    if (total_pairs > 1000000000ull) return set_err(c, HG_ERR_NOMEM, "hg_layout: more than 1e9 pairs between maximal reads");
    const size_t np_total = std::max<size_t>((size_t)total_pairs, 1), nc_slots = 2 * np_total;
    // bucket scratch: the count after the last insertion is below 2 n + 13 per read
    const size_t nbkt = 2 * np_total + 16 * (size_t)n + 64;
    HG_TRY(d_pairs.alloc(c, np_total, "pairs")); HG_TRY(d_hnext.alloc(c, np_total, "hash next"));
    HG_TRY(d_hout.alloc(c, np_total, "hash order")); HG_TRY(d_hbkt.alloc(c, nbkt, "hash buckets"));
    HG_TRY(d_sort.alloc(c, nc_slots, "sort scratch"));
    Cand* d_cands = nullptr;
    int* d_order = nullptr;
    int4* d_ranges = nullptr;
    HG_TRY(dev_alloc(c, &d_cands, nc_slots, "candidates"));
    R.d_cands = d_cands;
    HG_TRY(dev_alloc(c, &d_order, nc_slots, "candidate order"));
    R.d_order = d_order;
    HG_TRY(dev_alloc(c, &d_ranges, n, "candidate ranges"));
    R.d_ranges = d_ranges;
    R.n_cand_slots = (int64_t)nc_slots;

    io.n_read = n; io.active = d_active.p; io.cands = d_cands; io.ranges = d_ranges; io.ranges_out = d_ranges;
    io.order = d_order; io.sort_scratch = d_sort.p;
    io.hv.off = d_hoff.p; io.hv.pos = d_hpos.p; io.hv.type = d_htype.p;
    io.kv.off = d_koff.p; io.kv.pos = d_kpos.p; io.kv.type = d_ktype.p;
    io.nk.off = d_noff.p; io.nk.pos = nullptr; io.nk.type = nullptr;
    io.hinge_alive = d_alive.p; io.counters = d_cnt.p + 8; io.chosen = d_chosen.p;
    io.graph = nullptr; io.nkout = nullptr; io.skips = nullptr;

    int big_sort_cap = 1 << 16;
    for (int attempt = 0;; attempt++) {
        HG_TRY(d_bigsort.alloc(c, big_sort_cap, "pair sort scratch"));
        L.pair_ref = d_pair_ref.p; L.cand_ref = d_cand_ref.p; L.bkt_ref = d_bkt_ref.p; L.pairs = d_pairs.p;
        L.cands = d_cands; L.hash_next = d_hnext.p; L.hash_out = d_hout.p; L.hash_bkt = d_hbkt.p;
        L.contained_flag = d_contained.p; L.counters = d_cnt.p; L.sort_scratch = d_bigsort.p; L.sort_cap = big_sort_cap;
        L.grow_at = d_grow.p; L.grow_bkt = d_grow.p + ngrow; L.ngrow = ngrow;
        if (attempt == 0) seg_begin();
        launch_layout_pairs(c->rec_view(), c->read_view(), P, c->fs.mask, d_active.p, L, st);
        if (attempt == 0) seg_end();
        int cnt[8];
        HG_TRY(cuda_check(c, cudaGetLastError(), "layout pairs"));
        HG_TRY(cuda_check(c, cudaMemcpyAsync(cnt, d_cnt.p, sizeof cnt, cudaMemcpyDeviceToHost, st), "D2H"));
        if (contained_out) HG_TRY(cuda_check(c, cudaMemcpyAsync(contained_out, d_contained.p, n, cudaMemcpyDeviceToHost, st), "D2H"));
        HG_TRY(cuda_check(c, cudaStreamSynchronize(st), "layout pairs"));
        if (cnt[2] > (long long)nbkt) return set_err(c, HG_ERR_NOMEM, "hg_layout: hash bucket scratch too small");
        if (!cnt[3]) break;
        if (attempt > 6) return set_err(c, HG_ERR_NOMEM, "pair sort scratch kept overflowing");
        big_sort_cap = std::max(big_sort_cap * 4, cnt[4] + 1024);
        // the contained flags are sticky and the pair counts were overwritten: restore and run again
        cudaMemsetAsync(d_contained.p, 0, n, st);
        cudaMemsetAsync(d_pair_ref.p, 0, sizeof(int2) * (size_t)n, st);
        launch_layout_count_pairs(c->rec_view(), c->read_view(), d_active.p, d_pair_ref.p, d_total.p, st);
    }
    tm.lap("layout: pairs + candidates (device)");
    return HG_OK;
}

// contained_all: the "[contained]" flags of ALL reads (null: this context's own are all there is)
int LayoutRun::phase2(const uint8_t* contained_all, uint8_t* alive_out) {
    cudaStream_t st = c->stream;
    LayoutResult& R = *c->layout;
    StepTimer tm(st);
    if (contained_all) HG_TRY(h2d(d_contained.p, contained_all, n));
    // "[contained] Should not happen" (hinging.cpp:598-601): such reads leave the layout
    seg_begin();
    launch_apply_contained(n, d_contained.p, d_active.p, d_cnt.p + 7, st);
    launch_order_candidates(L, io, st);
    seg_end();
    HG_TRY(cuda_check(c, cudaMemcpyAsync(&R.n_contained, d_cnt.p + 7, 4, cudaMemcpyDeviceToHost, st), "D2H"));

    // ---- K6: kill pass + hinge graph; list sizes are data dependent: grow and rerun on overflow
    int graph_cap = 1 << 16, nk_cap = 1 << 14;
    for (int attempt = 0;; attempt++) {
        DevBuf<GraphRec> d_graph;
        DevBuf<NkRec> d_nk;
        HG_TRY(d_graph.alloc(c, graph_cap, "graph records"));
        HG_TRY(d_nk.alloc(c, nk_cap, "nk records"));
        io.graph = d_graph.p; io.graph_cap = graph_cap; io.nkout = d_nk.p; io.nk_cap = nk_cap;
        io.skips = nullptr; io.skip_cap = 0;
        cudaMemsetAsync(d_alive.p, 1, std::max<size_t>((size_t)nh, 1), st);
        if (attempt == 0) seg_begin();
        launch_hinge_graph(c->rec_view(), P, io, st);
        if (attempt == 0) seg_end();
        int cnt[8];
        HG_TRY(cuda_check(c, cudaMemcpyAsync(cnt, io.counters, sizeof cnt, cudaMemcpyDeviceToHost, st), "D2H"));
        HG_TRY(cuda_check(c, cudaStreamSynchronize(st), "hinge graph"));
        if (cnt[3]) {
            if (attempt > 8) return set_err(c, HG_ERR_NOMEM, "hinge graph buffers kept overflowing");
            graph_cap = std::max(graph_cap, cnt[0] + 1024);
            nk_cap = std::max(nk_cap, cnt[1] + 1024);
            continue;
        }
        graph.resize(cnt[0]);
        nks.resize(cnt[1]);
        if (cnt[0]) HG_TRY(cuda_check(c, cudaMemcpyAsync(graph.data(), d_graph.p, sizeof(GraphRec) * (size_t)cnt[0], cudaMemcpyDeviceToHost, st), "D2H"));
        if (cnt[1]) HG_TRY(cuda_check(c, cudaMemcpyAsync(nks.data(), d_nk.p, sizeof(NkRec) * (size_t)cnt[1], cudaMemcpyDeviceToHost, st), "D2H"));
        HG_TRY(cuda_check(c, cudaStreamSynchronize(st), "D2H"));
        break;
    }
    for (int i = 0; i < R.n_contained; i++) printf("[contained] Should not happen\n");
    R.hin_alive.assign((size_t)std::max<int64_t>(nh, 1), 1);
    if (nh) HG_TRY(cuda_check(c, cudaMemcpyAsync(R.hin_alive.data(), d_alive.p, (size_t)nh, cudaMemcpyDeviceToHost, st), "D2H"));
    HG_TRY(cuda_check(c, cudaStreamSynchronize(st), "D2H"));
    if (alive_out && nh) memcpy(alive_out, R.hin_alive.data(), (size_t)nh);
    tm.lap("layout: order + kill pass + hinge graph + D2H");
    return HG_OK;
}

// alive_all / graph_all: the kill-pass result and the hinge-graph records of ALL ranks (null: this
// context's own)
int LayoutRun::phase3(const uint8_t* alive_all, const GraphRec* graph_all, int64_t n_graph_all) {
    cudaStream_t st = c->stream;
    LayoutResult& R = *c->layout;
    StepTimer tm(st);
    if (graph_all) graph.assign(graph_all, graph_all + n_graph_all);
    if (alive_all && nh) memcpy(R.hin_alive.data(), alive_all, (size_t)nh);
    std::sort(graph.begin(), graph.end(), [](const GraphRec& x, const GraphRec& y) {
        return x.owner != y.owner ? x.owner < y.owner : x.seq < y.seq;
    });
    std::sort(nks.begin(), nks.end(), [](const NkRec& x, const NkRec& y) {
        return x.owner != y.owner ? x.owner < y.owner : x.seq < y.seq;
    });
    // connected components of the hinge graph (hinging.cpp:1644-1675): only sizes matter
    {
        std::vector<int> parent((size_t)nh), size((size_t)nh, 0);
        std::iota(parent.begin(), parent.end(), 0);
        auto find = [&](int x) {
            while (parent[x] != x) {
                parent[x] = parent[parent[x]];
                x = parent[x];
            }
            return x;
        };
        for (const GraphRec& g : graph)
            if (g.flag == 1) parent[find(g.u)] = find(g.v);
        for (int64_t v = 0; v < nh; v++) size[find((int)v)]++;
        for (int64_t v = 0; v < nh; v++)
            if (size[find((int)v)] < P.min_connected_component_size) R.hin_alive[v] = 0;
    }
    HG_TRY(h2d(d_alive.p, R.hin_alive.data(), (size_t)nh));
    // new_killed_hinges_vec as a CSR
    std::vector<int64_t> noff((size_t)n + 1, 0);
    std::vector<int> npos(nks.size()), ntype(nks.size());
    for (const NkRec& r : nks) noff[r.owner + 1]++;
    for (int i = 0; i < n; i++) noff[i + 1] += noff[i];
    for (size_t k = 0; k < nks.size(); k++) {
        npos[k] = nks[k].pos;
        ntype[k] = nks[k].type;
    }
    HG_TRY(d_npos.alloc(c, nks.size(), "nk pos"));
    HG_TRY(d_ntype.alloc(c, nks.size(), "nk type"));
    HG_TRY(h2d(d_noff.p, noff.data(), 8 * ((size_t)n + 1)));
    HG_TRY(h2d(d_npos.p, npos.data(), 4 * nks.size()));
    HG_TRY(h2d(d_ntype.p, ntype.data(), 4 * nks.size()));
    io.nk.pos = d_npos.p;
    io.nk.type = d_ntype.p;
    R.graph.swap(graph);
    tm.lap("layout: components + nk lists (host)");

    // ---- K6: the best-overlap scoring loop
    int skip_cap = 1 << 14;
    std::vector<SkipRec> skips;
    R.chosen.assign(2 * (size_t)n, make_int2(-1, -1));
    DevBuf<Cand> d_edges;
    DevBuf<int2> d_edge_ref;
    HG_TRY(d_edges.alloc(c, 2 * (size_t)n, "edges"));
    HG_TRY(d_edge_ref.alloc(c, 2 * (size_t)n, "edge refs"));
    for (int attempt = 0;; attempt++) {
        DevBuf<SkipRec> d_skip;
        HG_TRY(d_skip.alloc(c, skip_cap, "skip records"));
        io.skips = d_skip.p;
        io.skip_cap = skip_cap;
        cudaMemsetAsync(io.counters, 0, sizeof(int) * 8, st);
        if (attempt == 0) seg_begin();
        launch_best_extension(P, io, st);
        launch_gather_chosen(n, d_chosen.p, R.d_cands, d_edges.p, d_edge_ref.p, io.counters + 6, st);
        if (attempt == 0) seg_end();
        cudaEventRecord(c->ev1, st);
        int cnt[8];
        HG_TRY(cuda_check(c, cudaGetLastError(), "best extension"));
        HG_TRY(cuda_check(c, cudaMemcpyAsync(cnt, io.counters, sizeof cnt, cudaMemcpyDeviceToHost, st), "D2H"));
        HG_TRY(cuda_check(c, cudaStreamSynchronize(st), "best extension"));
        if (cnt[3]) {
            if (attempt > 8) return set_err(c, HG_ERR_NOMEM, "skip buffer kept overflowing");
            skip_cap = std::max(skip_cap, cnt[2] + 1024);
            continue;
        }
        skips.resize(cnt[2]);
        R.edge_cands.resize(cnt[6]);
        R.edge_ref.resize(cnt[6]);
        if (cnt[2]) HG_TRY(cuda_check(c, cudaMemcpyAsync(skips.data(), d_skip.p, sizeof(SkipRec) * (size_t)cnt[2], cudaMemcpyDeviceToHost, st), "D2H"));
        if (cnt[6]) {
            HG_TRY(cuda_check(c, cudaMemcpyAsync(R.edge_cands.data(), d_edges.p, sizeof(Cand) * (size_t)cnt[6], cudaMemcpyDeviceToHost, st), "D2H"));
            HG_TRY(cuda_check(c, cudaMemcpyAsync(R.edge_ref.data(), d_edge_ref.p, sizeof(int2) * (size_t)cnt[6], cudaMemcpyDeviceToHost, st), "D2H"));
        }
        HG_TRY(cuda_check(c, cudaMemcpyAsync(R.chosen.data(), d_chosen.p, sizeof(int2) * 2 * (size_t)n, cudaMemcpyDeviceToHost, st), "D2H"));
        HG_TRY(cuda_check(c, cudaMemcpyAsync(R.active.data(), d_active.p, n, cudaMemcpyDeviceToHost, st), "D2H"));
        HG_TRY(cuda_check(c, cudaStreamSynchronize(st), "D2H"));
        break;
    }
    std::sort(skips.begin(), skips.end(), [](const SkipRec& x, const SkipRec& y) {
        return x.owner != y.owner ? x.owner < y.owner : x.seq < y.seq;
    });
    R.skips.swap(skips);
    // edges in output order: by read, forward before backward
    {
        std::vector<int> idx(R.edge_ref.size());
        std::iota(idx.begin(), idx.end(), 0);
        std::sort(idx.begin(), idx.end(), [&](int x, int y) { return R.edge_ref[x].x < R.edge_ref[y].x; });
        std::vector<Cand> ec(idx.size());
        std::vector<int2> er(idx.size());
        for (size_t k = 0; k < idx.size(); k++) {
            ec[k] = R.edge_cands[idx[k]];
            er[k] = R.edge_ref[idx[k]];
        }
        R.edge_cands.swap(ec);
        R.edge_ref.swap(er);
    }
    tm.lap("layout: best extension + D2H");
    kernel_ms = kernel_time();
    return HG_OK;
}

}  // namespace hg

extern "C" {

static int layout_args_ok(hg_ctx* c, const hg_layout_params* P, const int32_t* mask, const uint8_t* maximal,
                          const int64_t* rep_off, const int64_t* hin_off) {
    if (!c || !P || !mask || !maximal || !rep_off || !hin_off || c->novl <= 0)
        return set_err(c, HG_ERR_ARG, "hg_layout: bad arguments");
    if (!c->has_trace) return set_err(c, HG_ERR_ARG, "hg_layout needs the trace (pass trace_off/trace to hg_set_overlaps)");
    return HG_OK;
}

int hg_layout(hg_ctx* c, const hg_layout_params* P, const int32_t* mask, const uint8_t* maximal,
              const int64_t* rep_off, const int32_t* rep_pos, const int32_t* rep_type,
              const int64_t* hin_off, const int32_t* hin_pos, const int32_t* hin_type,
              float* ms_device) {
    HG_TRY(layout_args_ok(c, P, mask, maximal, rep_off, hin_off));
    if (c->a_lo != 0 || c->a_hi != c->n_read)
        return set_err(c, HG_ERR_ARG, "hg_layout on a context that owns a slice of the reads: use hg_layout_phase1..3 "
                                      "and exchange the flags between them");
    cudaSetDevice(c->device);
    free_layout_run(c->layout_run);
    c->layout_run = new LayoutRun();
    c->layout_run->c = c;
    int rc = c->layout_run->phase1(P, mask, maximal, rep_off, rep_pos, rep_type, hin_off, hin_pos, hin_type, nullptr);
    if (rc == HG_OK) rc = c->layout_run->phase2(nullptr, nullptr);
    if (rc == HG_OK) rc = c->layout_run->phase3(nullptr, nullptr, 0);
    if (rc == HG_OK && ms_device) *ms_device = c->layout_run->kernel_ms;
    free_layout_run(c->layout_run);  // the candidate lists live on in c->layout
    c->layout_run = nullptr;
    return rc;
}

// ---- sharded form: every rank classifies the pairs of its own reads and picks their edges; what
// crosses shards is small and travels as host arrays through the caller's collectives (hinge_b200/
// sharding.py, NCCL):
//   phase 1 -> contained_out[n_read]  "[contained]" flags of the own reads      : MAX all-reduce
//   phase 2 -> alive_out[n_hinges]    hinges the own reads' matches killed (0)  : MIN all-reduce
//              graph records (hg_layout_graph)                                  : all-gather
//   phase 3    components of the whole hinge graph, best extension of the own reads
int hg_layout_phase1(hg_ctx* c, const hg_layout_params* P, const int32_t* mask, const uint8_t* maximal,
                     const int64_t* rep_off, const int32_t* rep_pos, const int32_t* rep_type,
                     const int64_t* hin_off, const int32_t* hin_pos, const int32_t* hin_type,
                     uint8_t* contained_out) {
    HG_TRY(layout_args_ok(c, P, mask, maximal, rep_off, hin_off));
    if (!contained_out) return set_err(c, HG_ERR_ARG, "hg_layout_phase1: contained_out is null");
    cudaSetDevice(c->device);
    free_layout_run(c->layout_run);
    c->layout_run = new LayoutRun();
    c->layout_run->c = c;
    return c->layout_run->phase1(P, mask, maximal, rep_off, rep_pos, rep_type, hin_off, hin_pos, hin_type, contained_out);
}

int hg_layout_phase2(hg_ctx* c, const uint8_t* contained_all, uint8_t* alive_out, int64_t* n_graph) {
    if (!c || !c->layout_run || !contained_all || !n_graph) return set_err(c, HG_ERR_ARG, "hg_layout_phase2 before phase1");
    cudaSetDevice(c->device);
    HG_TRY(c->layout_run->phase2(contained_all, alive_out));
    *n_graph = (int64_t)c->layout_run->graph.size();
    return HG_OK;
}

int hg_layout_graph(hg_ctx* c, hg_graph_rec* out, int64_t capacity) {
    if (!c || !c->layout_run || (!out && capacity > 0)) return set_err(c, HG_ERR_ARG, "hg_layout_graph before phase2");
    static_assert(sizeof(hg_graph_rec) == sizeof(GraphRec), "graph record layout");
    const int64_t m = std::min<int64_t>(capacity, (int64_t)c->layout_run->graph.size());
    if (m > 0) memcpy(out, c->layout_run->graph.data(), sizeof(GraphRec) * (size_t)m);
    return HG_OK;
}

int hg_layout_phase3(hg_ctx* c, const uint8_t* alive_all, const hg_graph_rec* graph_all, int64_t n_graph_all,
                     float* ms_device) {
    if (!c || !c->layout_run || !alive_all || (!graph_all && n_graph_all > 0))
        return set_err(c, HG_ERR_ARG, "hg_layout_phase3 before phase2");
    cudaSetDevice(c->device);
    const int rc = c->layout_run->phase3(alive_all, reinterpret_cast<const GraphRec*>(graph_all), n_graph_all);
    if (rc == HG_OK && ms_device) *ms_device = c->layout_run->kernel_ms;
    free_layout_run(c->layout_run);
    c->layout_run = nullptr;
    return rc;
}

int hg_layout_edges(hg_ctx* c, hg_edge* edges, int64_t capacity, int64_t* n_edges) {
    if (!c || !c->layout || !n_edges) return set_err(c, HG_ERR_ARG, "hg_layout_edges: run hg_layout first");
    const LayoutResult& R = *c->layout;
    const int64_t n = (int64_t)R.edge_cands.size();
    for (int64_t k = 0; k < n && edges && k < capacity; k++) R.fill_edge(R.edge_cands[k], R.edge_ref[k].y, &edges[k]);
    *n_edges = n;
    return HG_OK;
}

}  // extern "C"
