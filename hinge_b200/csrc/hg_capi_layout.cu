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

const LayoutResult* layout_result(const hg_ctx* c) { return c->layout; }
void free_layout_result(LayoutResult* r) { delete r; }

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

int hg_layout(hg_ctx* c, const hg_layout_params* P, const int32_t* mask, const uint8_t* maximal,
              const int64_t* rep_off, const int32_t* rep_pos, const int32_t* rep_type,
              const int64_t* hin_off, const int32_t* hin_pos, const int32_t* hin_type,
              float* ms_device) {
    if (!c || !P || !mask || !maximal || !rep_off || !hin_off || c->novl <= 0)
        return set_err(c, HG_ERR_ARG, "hg_layout: bad arguments");
    if (!c->has_trace) return set_err(c, HG_ERR_ARG, "hg_layout needs the trace (pass trace_off/trace to hg_set_overlaps)");
    if (c->a_lo != 0 || c->a_hi != c->n_read)
        return set_err(c, HG_ERR_ARG, "hg_layout runs on a context that owns all reads (gather the shards' results first)");
    cudaSetDevice(c->device);
    cudaStream_t st = c->stream;
    const int n = c->n_read;
    free_layout_result(c->layout);
    c->layout = new LayoutResult();
    LayoutResult& R = *c->layout;
    R.n_read = n;
    R.mask.assign(mask, mask + 2 * (size_t)n);

    // ---- who takes part (hinging.cpp:877-913, 954-960, 398-412)
    R.active.assign(n, 1);
    for (int i = 0; i < n; i++) {
        if (P->delete_telomeres && rep_off[i + 1] - rep_off[i] > P->num_events_telomere) R.active[i] = 0;
        if (mask[2 * i + 1] - mask[2 * i] < P->length_threshold) {
            R.active[i] = 0;
            R.garbage.push_back(i);
        }
        R.active[i] = R.active[i] && maximal[i];
    }
    HG_TRY(upload_mask(c, mask));
    DevBuf<uint8_t> d_active;
    HG_TRY(d_active.alloc(c, n, "active"));
    HG_TRY(cuda_check(c, cudaMemcpyAsync(d_active.p, R.active.data(), n, cudaMemcpyHostToDevice, st), "H2D"));

    // ---- K5: classify the top two overlaps of every pair of maximal reads
    StepTimer tm(st);
    tm.lap("layout: active flags + uploads");
    cudaEventRecord(c->ev0, st);
    int big_cap = 1 << 14, sort_cap = 1 << 18, pair_cap = 1 << 20, cand_cap = 1 << 20;
    std::vector<int4> pairs;
    std::vector<Cand> cands;
    std::vector<uint8_t> contained(n);
    for (int attempt = 0;; attempt++) {
        PairBufs pb;
        HG_TRY(pb.make(c, big_cap, sort_cap, pair_cap, cand_cap));
        cudaMemsetAsync(pb.po.contained_flag, 0, n, st);
        launch_classify(c->rec_view(), c->read_view(), *P, c->fs.mask, d_active.p, 1, 1, nullptr, pb.po, st);
        int cnt[8];
        HG_TRY(cuda_check(c, cudaMemcpyAsync(cnt, pb.po.counters, sizeof cnt, cudaMemcpyDeviceToHost, st), "D2H"));
        HG_TRY(cuda_check(c, cudaStreamSynchronize(st), "classify"));
        if (cnt[3]) {
            if (attempt > 6) return set_err(c, HG_ERR_NOMEM, "candidate buffers kept overflowing");
            big_cap = std::max(big_cap, cnt[0] + 1024);
            sort_cap = std::max(sort_cap * 4, cnt[1] + 1024);
            pair_cap = std::max(pair_cap, cnt[4] + 1024);
            cand_cap = std::max(cand_cap, cnt[5] + 1024);
            continue;
        }
        pairs.resize(cnt[4]);
        cands.resize(cnt[5]);
        if (cnt[4]) HG_TRY(cuda_check(c, cudaMemcpyAsync(pairs.data(), pb.po.pairs, sizeof(int4) * (size_t)cnt[4], cudaMemcpyDeviceToHost, st), "D2H"));
        if (cnt[5]) HG_TRY(cuda_check(c, cudaMemcpyAsync(cands.data(), pb.po.cands, sizeof(Cand) * (size_t)cnt[5], cudaMemcpyDeviceToHost, st), "D2H"));
        HG_TRY(cuda_check(c, cudaMemcpyAsync(contained.data(), pb.po.contained_flag, n, cudaMemcpyDeviceToHost, st), "D2H"));
        HG_TRY(cuda_check(c, cudaStreamSynchronize(st), "D2H"));
        break;
    }
    tm.lap("layout: classify pairs + D2H");
    for (int i = 0; i < n; i++)
        if (contained[i] && R.active[i]) {  // hinging.cpp:598-601
            printf("[contained] Should not happen\n");
            R.active[i] = 0;
        }
    HG_TRY(cuda_check(c, cudaMemcpyAsync(d_active.p, R.active.data(), n, cudaMemcpyHostToDevice, st), "H2D"));

    // ---- pre-sort order of each read's candidates = iteration order of the reference's
    // std::unordered_map<int, ...> keyed by B (hinging.cpp:532): replay the keys, in file order,
    // into the real container
    auto first_of = [](const int4& p) { return ((int64_t)p.w << 31) | (int64_t)p.z; };
    std::sort(pairs.begin(), pairs.end(), [&](const int4& x, const int4& y) { return first_of(x) < first_of(y); });
    std::sort(cands.begin(), cands.end(), [](const Cand& x, const Cand& y) {
        if (x.a != y.a) return x.a < y.a;
        if (x.b != y.b) return x.b < y.b;
        return x.rank < y.rank;
    });
    R.cands = cands;
    R.ranges.assign(n, make_int4(0, 0, 0, 0));
    R.order.clear();
    {
        size_t pi = 0, ci = 0;
        for (int a = 0; a < n; a++) {
            const size_t p0 = pi;
            while (pi < pairs.size() && pairs[pi].x == a) pi++;
            const size_t c0 = ci;
            while (ci < cands.size() && cands[ci].a == a) ci++;
            if (p0 == pi) continue;
            std::unordered_map<int, int> um;
            for (size_t k = p0; k < pi; k++) um[pairs[k].y] = 0;
            std::vector<int> fwd, bwd;
            for (auto it = um.begin(); it != um.end(); ++it) {
                const int b = it->first;
                size_t lo = std::lower_bound(cands.begin() + c0, cands.begin() + ci, b,
                                             [](const Cand& x, int v) { return x.b < v; }) - cands.begin();
                for (size_t k = lo; k < ci && cands[k].b == b; k++) {
                    const bool f = cands[k].type == HG_FORWARD || cands[k].type == HG_FORWARD_INTERNAL;
                    (f ? fwd : bwd).push_back((int)k);
                }
            }
            int4 r;
            r.x = (int)R.order.size();
            R.order.insert(R.order.end(), fwd.begin(), fwd.end());
            r.y = (int)R.order.size();
            r.z = r.y;
            R.order.insert(R.order.end(), bwd.begin(), bwd.end());
            r.w = (int)R.order.size();
            R.ranges[a] = r;
        }
    }

    tm.lap("layout: host sorts + hash-order replay");
    // ---- hinges, killed hinges (hinging.cpp:1180-1197)
    R.hin_off.assign(hin_off, hin_off + n + 1);
    R.hin_pos.assign(hin_pos, hin_pos + hin_off[n]);
    R.hin_type.assign(hin_type, hin_type + hin_off[n]);
    R.kil_off.assign((size_t)n + 1, 0);
    R.kil_pos.clear();
    R.kil_type.clear();
    for (int i = 0; i < n; i++) {
        std::set<std::pair<int, int>> surviving;
        for (int64_t k = hin_off[i]; k < hin_off[i + 1]; k++) surviving.insert(std::make_pair(hin_pos[k], hin_type[k]));
        for (int64_t k = rep_off[i]; k < rep_off[i + 1]; k++)
            if (!surviving.count(std::make_pair(rep_pos[k], rep_type[k]))) {
                R.kil_pos.push_back(rep_pos[k]);
                R.kil_type.push_back(rep_type[k]);
            }
        R.kil_off[i + 1] = (int64_t)R.kil_pos.size();
    }
    const int64_t nh = hin_off[n], nkil = R.kil_off[n];
    tm.lap("layout: killed-hinge lists (host)");

    // ---- device side of the selection
    const size_t ncand = std::max<size_t>(cands.size(), 1), nord = std::max<size_t>(R.order.size(), 1);
    DevBuf<Cand> d_cands;
    DevBuf<int4> d_ranges;
    DevBuf<int> d_order, d_hpos, d_htype, d_kpos, d_ktype, d_npos, d_ntype, d_cnt;
    DevBuf<int64_t> d_hoff, d_koff, d_noff;
    DevBuf<KeyIdx2> d_sort;
    DevBuf<uint8_t> d_alive;
    DevBuf<int2> d_chosen;
    HG_TRY(d_cands.alloc(c, ncand, "cands")); HG_TRY(d_ranges.alloc(c, n, "ranges"));
    HG_TRY(d_order.alloc(c, nord, "order")); HG_TRY(d_sort.alloc(c, nord, "sort scratch"));
    HG_TRY(d_hoff.alloc(c, n + 1, "hinge off")); HG_TRY(d_hpos.alloc(c, nh, "hinge pos")); HG_TRY(d_htype.alloc(c, nh, "hinge type"));
    HG_TRY(d_koff.alloc(c, n + 1, "killed off")); HG_TRY(d_kpos.alloc(c, nkil, "killed pos")); HG_TRY(d_ktype.alloc(c, nkil, "killed type"));
    HG_TRY(d_noff.alloc(c, n + 1, "nk off"));
    HG_TRY(d_alive.alloc(c, nh, "hinge alive")); HG_TRY(d_cnt.alloc(c, 8, "counters")); HG_TRY(d_chosen.alloc(c, 2 * (size_t)n, "chosen"));
    auto h2d = [&](void* d, const void* h, size_t bytes) {
        return bytes ? cuda_check(c, cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st), "H2D") : HG_OK;
    };
    HG_TRY(h2d(d_cands.p, cands.data(), sizeof(Cand) * cands.size()));
    HG_TRY(h2d(d_ranges.p, R.ranges.data(), sizeof(int4) * n));
    HG_TRY(h2d(d_order.p, R.order.data(), 4 * R.order.size()));
    HG_TRY(h2d(d_hoff.p, R.hin_off.data(), 8 * ((size_t)n + 1)));
    HG_TRY(h2d(d_hpos.p, R.hin_pos.data(), 4 * (size_t)nh));
    HG_TRY(h2d(d_htype.p, R.hin_type.data(), 4 * (size_t)nh));
    HG_TRY(h2d(d_koff.p, R.kil_off.data(), 8 * ((size_t)n + 1)));
    HG_TRY(h2d(d_kpos.p, R.kil_pos.data(), 4 * (size_t)nkil));
    HG_TRY(h2d(d_ktype.p, R.kil_type.data(), 4 * (size_t)nkil));
    cudaMemsetAsync(d_alive.p, 1, std::max<size_t>((size_t)nh, 1), st);

    SelectIO io;
    io.n_read = n; io.active = d_active.p; io.cands = d_cands.p; io.ranges = d_ranges.p;
    io.order = d_order.p; io.sort_scratch = d_sort.p;
    io.hv.off = d_hoff.p; io.hv.pos = d_hpos.p; io.hv.type = d_htype.p;
    io.kv.off = d_koff.p; io.kv.pos = d_kpos.p; io.kv.type = d_ktype.p;
    io.nk.off = d_noff.p; io.nk.pos = nullptr; io.nk.type = nullptr;
    io.hinge_alive = d_alive.p; io.counters = d_cnt.p; io.chosen = d_chosen.p;
    io.graph = nullptr; io.nkout = nullptr; io.skips = nullptr;

    tm.lap("layout: selection uploads");
    launch_sort_candidates(io, st);  // K6: weight order (hinging.cpp:1066-1071)
    tm.lap("layout: sort candidates");

    // K6: kill pass + hinge graph; list sizes are data dependent: grow and rerun on overflow
    int graph_cap = 1 << 16, nk_cap = 1 << 14;
    std::vector<GraphRec> graph;
    std::vector<NkRec> nks;
    for (int attempt = 0;; attempt++) {
        DevBuf<GraphRec> d_graph;
        DevBuf<NkRec> d_nk;
        HG_TRY(d_graph.alloc(c, graph_cap, "graph records"));
        HG_TRY(d_nk.alloc(c, nk_cap, "nk records"));
        io.graph = d_graph.p; io.graph_cap = graph_cap; io.nkout = d_nk.p; io.nk_cap = nk_cap;
        io.skips = nullptr; io.skip_cap = 0;
        cudaMemsetAsync(d_alive.p, 1, std::max<size_t>((size_t)nh, 1), st);
        launch_hinge_graph(c->rec_view(), *P, io, st);
        int cnt[8];
        HG_TRY(cuda_check(c, cudaMemcpyAsync(cnt, d_cnt.p, sizeof cnt, cudaMemcpyDeviceToHost, st), "D2H"));
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
    tm.lap("layout: kill pass + hinge graph + D2H");
    std::sort(graph.begin(), graph.end(), [](const GraphRec& x, const GraphRec& y) {
        return x.owner != y.owner ? x.owner < y.owner : x.seq < y.seq;
    });
    std::sort(nks.begin(), nks.end(), [](const NkRec& x, const NkRec& y) {
        return x.owner != y.owner ? x.owner < y.owner : x.seq < y.seq;
    });
    R.graph = graph;

    // connected components of the hinge graph (hinging.cpp:1644-1675): only sizes matter
    R.hin_alive.assign((size_t)std::max<int64_t>(nh, 1), 1);
    if (nh) HG_TRY(cuda_check(c, cudaMemcpyAsync(R.hin_alive.data(), d_alive.p, (size_t)nh, cudaMemcpyDeviceToHost, st), "D2H"));
    HG_TRY(cuda_check(c, cudaStreamSynchronize(st), "D2H"));
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
            if (size[find((int)v)] < P->min_connected_component_size) R.hin_alive[v] = 0;
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

    tm.lap("layout: components + nk lists (host)");
    // K6: the best-overlap scoring loop
    int skip_cap = 1 << 14;
    std::vector<SkipRec> skips;
    R.chosen.assign(2 * (size_t)n, make_int2(-1, -1));
    for (int attempt = 0;; attempt++) {
        DevBuf<SkipRec> d_skip;
        HG_TRY(d_skip.alloc(c, skip_cap, "skip records"));
        io.skips = d_skip.p;
        io.skip_cap = skip_cap;
        cudaMemsetAsync(d_cnt.p, 0, sizeof(int) * 8, st);
        launch_best_extension(*P, io, st);
        int cnt[8];
        HG_TRY(cuda_check(c, cudaMemcpyAsync(cnt, d_cnt.p, sizeof cnt, cudaMemcpyDeviceToHost, st), "D2H"));
        HG_TRY(cuda_check(c, cudaStreamSynchronize(st), "best extension"));
        if (cnt[3]) {
            if (attempt > 8) return set_err(c, HG_ERR_NOMEM, "skip buffer kept overflowing");
            skip_cap = std::max(skip_cap, cnt[2] + 1024);
            continue;
        }
        skips.resize(cnt[2]);
        if (cnt[2]) HG_TRY(cuda_check(c, cudaMemcpyAsync(skips.data(), d_skip.p, sizeof(SkipRec) * (size_t)cnt[2], cudaMemcpyDeviceToHost, st), "D2H"));
        HG_TRY(cuda_check(c, cudaMemcpyAsync(R.chosen.data(), d_chosen.p, sizeof(int2) * 2 * (size_t)n, cudaMemcpyDeviceToHost, st), "D2H"));
        if (!R.order.empty()) HG_TRY(cuda_check(c, cudaMemcpyAsync(R.order.data(), d_order.p, 4 * R.order.size(), cudaMemcpyDeviceToHost, st), "D2H"));
        cudaEventRecord(c->ev1, st);
        HG_TRY(cuda_check(c, cudaStreamSynchronize(st), "D2H"));
        break;
    }
    tm.lap("layout: best extension + D2H");
    std::sort(skips.begin(), skips.end(), [](const SkipRec& x, const SkipRec& y) {
        return x.owner != y.owner ? x.owner < y.owner : x.seq < y.seq;
    });
    R.skips = skips;
    if (ms_device) cudaEventElapsedTime(ms_device, c->ev0, c->ev1);
    return HG_OK;
}

int hg_layout_edges(hg_ctx* c, hg_edge* edges, int64_t capacity, int64_t* n_edges) {
    if (!c || !c->layout || !n_edges) return set_err(c, HG_ERR_ARG, "hg_layout_edges: run hg_layout first");
    const LayoutResult& R = *c->layout;
    int64_t k = 0;
    for (int i = 0; i < R.n_read; i++)
        for (int half = 0; half < 2; half++) {
            const int2 ch = R.chosen[2 * (size_t)i + half];
            if (ch.x < 0) continue;
            if (edges && k < capacity) R.fill_edge(ch.x, ch.y, &edges[k]);
            k++;
        }
    *n_edges = k;
    return HG_OK;
}

}  // extern "C"
