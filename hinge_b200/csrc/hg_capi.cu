// C ABI (include/hinge_b200.h): context, ingest, the filter stage.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "hg_ctx.h"

using namespace hg;

namespace hg {

int set_err(hg_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}

int cuda_check(hg_ctx* c, cudaError_t e, const char* what) {
    if (e == cudaSuccess) return HG_OK;
    return set_err(c, e == cudaErrorMemoryAllocation ? HG_ERR_NOMEM : HG_ERR_CUDA,
                   std::string(what) + ": " + cudaGetErrorString(e));
}

}  // namespace hg

RecView hg_ctx::rec_view() const {
    RecView v;
    v.novl = novl;
    v.aread = d_aread; v.bread = d_bread; v.abpos = d_abpos; v.aepos = d_aepos;
    v.bbpos = d_bbpos; v.bepos = d_bepos; v.flags = d_flags;
    v.trace_off = d_trace_off; v.trace = d_trace; v.tbytes = tbytes;
    v.read_off = d_read_off;
    return v;
}

ReadView hg_ctx::read_view() const {
    ReadView v;
    v.n_read = n_read;
    v.r_lo = a_lo;
    v.r_hi = a_hi;
    v.rlen = d_rlen;
    v.qvmask = d_qvmask;
    return v;
}

static void free_overlaps(hg_ctx* c) {
    c->cap_novl = 0;
    c->cap_trace = 0;
    if (!c->adopted) {
        cudaFree(c->d_aread); cudaFree(c->d_bread); cudaFree(c->d_abpos); cudaFree(c->d_aepos);
        cudaFree(c->d_bbpos); cudaFree(c->d_bepos); cudaFree(c->d_flags);
        cudaFree(c->d_trace_off); cudaFree(c->d_trace);
    }
    c->d_aread = c->d_bread = c->d_abpos = c->d_aepos = c->d_bbpos = c->d_bepos = c->d_flags = nullptr;
    c->d_trace_off = nullptr;
    c->d_trace = nullptr;
    c->novl = 0;
    c->has_trace = false;
}

extern "C" {

const char* hg_version(void) { return "hinge_b200 0.1 (sm_100a)"; }

const char* hg_last_error(const hg_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int hg_ctx_create(int device, void* stream, hg_ctx** out) {
    if (!out) return HG_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        fprintf(stderr, "hinge_b200: no usable CUDA device (%s); there is no CPU fallback\n",
                e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range");
        return HG_ERR_CUDA;
    }
    hg_ctx* c = new hg_ctx();
    c->device = device;
    c->stream = (cudaStream_t)stream;
    if (cudaSetDevice(device) != cudaSuccess) {
        delete c;
        return HG_ERR_CUDA;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->num_sms = prop.multiProcessorCount;
    c->fs.num_sms = c->num_sms;
    // tuning aid for the executables (the ABI has HG_OPT_PROFILE_KERNEL): form of the coverage-profile kernel
    if (const char* v = getenv("HINGE_B200_PROFILE_KERNEL")) {
        const int k = atoi(v);
        if (k == 0 || k == 1 || k == 3 || k == 5 || k == 6) c->fs.flat_kernel = k;
    }
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    for (int i = 0; i < 3; i++) cudaStreamCreateWithFlags(&c->fs.side_stream[i], cudaStreamNonBlocking);
    for (int i = 0; i < 4; i++) cudaEventCreateWithFlags(&c->fs.side_event[i], cudaEventDisableTiming);
    int rc = dev_alloc(c, &c->d_err, 4, "err flag");
    if (rc == HG_OK) rc = dev_alloc(c, &c->fs.scal, 8, "scalars");
    if (rc == HG_OK) rc = dev_alloc(c, &c->fs.counters, 16, "counters");
    if (rc == HG_OK) c->fs.counters1 = c->fs.counters + 8;
    if (rc == HG_OK) rc = dev_alloc(c, &c->fs.med_hist, 4098, "median histogram");
    if (rc == HG_OK) rc = cuda_check(c, cudaMemset(c->fs.med_hist, 0, sizeof(unsigned int) * 4098), "median histogram");
    if (rc != HG_OK) {
        hg_ctx_destroy(c);
        return rc;
    }
    *out = c;
    return HG_OK;
}

static void peer_teardown(hg_ctx* c) {
    for (int r = 0; r < kMaxPeers; r++) {
        if (c->peer_ipc[r] && c->peer.base[r]) cudaIpcCloseMemHandle(c->peer.base[r]);
        c->peer_ipc[r] = false;
        c->peer.base[r] = nullptr;
    }
    if (c->peer_block) {
        if (c->fs.mask_pk == reinterpret_cast<uint32_t*>(c->peer_block + kPeerFlagBytes + kPeerHistBytes))
            c->fs.mask_pk = nullptr;
        cudaFree(c->peer_block);
    }
    c->peer_block = nullptr;
    c->peer.world = 0;
    c->peer_connected = false;
}

void hg_ctx_destroy(hg_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    peer_teardown(c);
    free_overlaps(c);
    cudaFree(c->d_rlen); cudaFree(c->d_qvmask); cudaFree(c->d_read_off); cudaFree(c->d_err);
    FilterScratch& s = c->fs;
    cudaFree(s.cov_maxbin); cudaFree(s.self_cnt); cudaFree(s.flat_prof);
    cudaFree(s.flat_rbatch); cudaFree(s.flat_desc); cudaFree(s.flat_zmap); cudaFree(s.flat_cmap);
    if (!c->ext_mean_cov) cudaFree(s.mean_cov);
    if (!c->ext_mask) cudaFree(s.mask);
    for (int i = 0; i < hg_ctx::kMarks; i++)
        if (c->marks[i]) cudaEventDestroy(c->marks[i]);
    if (!c->ext_med_hist) cudaFree(s.med_hist);
    cudaFree(s.scal); cudaFree(s.cmask); cudaFree(s.rflags);
    cudaFree(s.anno_ref); cudaFree(s.anno_pool); cudaFree(s.counters); cudaFree(s.work_items);
    cudaFree(s.big_list); cudaFree(s.exact_list); cudaFree(s.big_scratch); cudaFree(s.hinge_keep); cudaFree(s.hinge_scratch);
    cudaFree(s.item_log); cudaFree(s.flat_batch); cudaFree(s.flat_rbase);
    cudaFree(c->d_cov0); cudaFree(c->d_cov0_off);
    cudaFree(c->ms.active0); cudaFree(c->ms.state); cudaFree(c->ms.rtype); cudaFree(c->ms.unk); cudaFree(c->ms.pool);
    cudaFree(c->ms.counters); cudaFree(c->ms.big_pairs); cudaFree(c->ms.sort_scratch);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (int i = 0; i < 3; i++)
        if (c->fs.side_stream[i]) cudaStreamDestroy(c->fs.side_stream[i]);
    for (int i = 0; i < 4; i++)
        if (c->fs.side_event[i]) cudaEventDestroy(c->fs.side_event[i]);
    free_layout_run(c->layout_run);
    free_layout_result(c->layout);
    delete c;
}

int hg_set_option(hg_ctx* c, int option, int64_t value) {
    if (!c) return HG_ERR_ARG;
    if (option == HG_OPT_KEEP_COVERAGE) {
        c->keep_cov = value != 0;
        return HG_OK;
    }
    if (option == HG_OPT_PROFILE) {
        c->profile = value != 0;
        if (c->profile && !c->marks[0])
            for (int i = 0; i < hg_ctx::kMarks; i++) cudaEventCreate(&c->marks[i]);
        if (c->profile && c->n_read > 0 && !c->fs.item_log)
            HG_TRY(dev_alloc(c, &c->fs.item_log, c->n_read, "item log"));
        return HG_OK;
    }
    if (option == HG_OPT_SCATTER_SPREAD) {
        if (value != 1 && value != 4 && value != 8 && value != 16)
            return set_err(c, HG_ERR_ARG, "HG_OPT_SCATTER_SPREAD: 1, 4, 8 or 16");
        c->fs.flat_spread = (int)value;
        return HG_OK;
    }
    if (option == HG_OPT_ANNO_POOL) {
        if (value < 1 || value > (1ll << 30)) return set_err(c, HG_ERR_ARG, "HG_OPT_ANNO_POOL: 1 .. 2^30 entries");
        c->anno_pool_hint = (int)value;
        c->fs.anno_cap = 0;  // reallocated by the next hg_set_overlaps
        return HG_OK;
    }
    if (option == HG_OPT_KEEP_MASKS) {
        c->keep_masks = value != 0;
        return HG_OK;
    }
    if (option == HG_OPT_PROFILE_KERNEL) {
        if (value != 0 && value != 1 && value != 3 && value != 5 && value != 6)
            return set_err(c, HG_ERR_ARG, "HG_OPT_PROFILE_KERNEL: 0, 1, 3, 5 or 6");
        c->fs.flat_kernel = (int)value;
        return HG_OK;
    }
    return set_err(c, HG_ERR_ARG, "unknown option");
}

int hg_set_reads(hg_ctx* c, int32_t n_read, const int32_t* rlen, const int64_t* qv_off,
                 const uint8_t* qv, int32_t tspace) {
    if (!c || n_read <= 0 || !rlen) return set_err(c, HG_ERR_ARG, "hg_set_reads: bad arguments");
    cudaSetDevice(c->device);
    cudaStream_t st = c->stream;
    if (c->peer_block) peer_teardown(c);  // sized by n_read
    c->g_begin = c->g_end = -1;
    c->n_read = n_read;
    c->tspace = tspace;
    c->h_rlen.assign(rlen, rlen + n_read);
    {
        std::vector<int> srt(c->h_rlen);
        std::sort(srt.begin(), srt.end());
        c->max_rlen = srt.back();
        c->rlen_q999 = srt[(size_t)((double)(n_read - 1) * 0.999)];
        if (srt.front() < 0) return set_err(c, HG_ERR_INPUT, "negative read length");
    }
    HG_TRY(dev_alloc(c, &c->d_rlen, (size_t)n_read + 8, "rlen"));  // + slack: bulk copies of 16-byte multiples (k_profile_tma)
    HG_TRY(dev_alloc(c, &c->d_qvmask, n_read, "qv mask"));
    HG_TRY(cuda_check(c, cudaMemcpyAsync(c->d_rlen, rlen, sizeof(int) * n_read, cudaMemcpyHostToDevice, st), "rlen H2D"));
    c->has_qv = qv_off != nullptr && qv != nullptr;
    if (c->has_qv) {
        int64_t* d_off = nullptr;
        uint8_t* d_qv = nullptr;
        const int64_t nq = qv_off[n_read];
        HG_TRY(dev_alloc(c, &d_off, (size_t)n_read + 1, "qv offsets"));
        int rc = dev_alloc(c, &d_qv, (size_t)nq, "qv");
        if (rc == HG_OK) rc = cuda_check(c, cudaMemcpyAsync(d_off, qv_off, 8 * ((size_t)n_read + 1), cudaMemcpyHostToDevice, st), "qv_off H2D");
        if (rc == HG_OK && nq > 0) rc = cuda_check(c, cudaMemcpyAsync(d_qv, qv, (size_t)nq, cudaMemcpyHostToDevice, st), "qv H2D");
        if (rc == HG_OK) {
            launch_qv_mask(n_read, d_off, d_qv, tspace, c->d_qvmask, st);
            rc = cuda_check(c, cudaStreamSynchronize(st), "qv mask");
        }
        cudaFree(d_off);
        cudaFree(d_qv);
        HG_TRY(rc);
    } else {
        HG_TRY(cuda_check(c, cudaMemsetAsync(c->d_qvmask, 0, sizeof(int2) * n_read, st), "qv mask"));
    }
    // per-read buffers
    FilterScratch& s = c->fs;
    HG_TRY(dev_alloc(c, &s.cov_maxbin, n_read, "cov_maxbin"));
    HG_TRY(dev_alloc(c, &s.self_cnt, n_read, "self_cnt"));
    if (c->ext_mean_cov) { s.mean_cov = nullptr; c->ext_mean_cov = false; }
    if (c->ext_mask) { s.mask = nullptr; c->ext_mask = false; }
    HG_TRY(dev_alloc(c, &s.mean_cov, n_read, "mean_cov"));
    HG_TRY(dev_alloc(c, &s.mask, n_read, "mask"));
    HG_TRY(dev_alloc(c, &s.cmask, n_read, "cmask"));
    HG_TRY(dev_alloc(c, &s.rflags, n_read, "rflags"));
    HG_TRY(dev_alloc(c, &s.anno_ref, n_read, "anno_ref"));
    HG_TRY(dev_alloc(c, &s.work_items, 3 * (size_t)n_read, "work items"));
    HG_TRY(dev_alloc(c, &s.big_list, n_read, "big_list"));
    HG_TRY(dev_alloc(c, &s.exact_list, n_read, "exact_list"));
    HG_TRY(dev_alloc(c, &c->d_read_off, (size_t)n_read + 1 + 4, "read_off"));  // + slack, as above
    cudaMemsetAsync(s.mean_cov, 0xff, sizeof(int) * n_read, st);
    cudaMemsetAsync(s.rflags, 0, n_read, st);
    s.anno_cap = 0;
    c->a_lo = 0;
    c->a_hi = n_read;
    c->filter_done = false;
    c->shape_version++;
    c->reads_version++;
    return cuda_check(c, cudaStreamSynchronize(st), "hg_set_reads");
}

static int alloc_anno_pool(hg_ctx* c, int cap) {
    FilterScratch& s = c->fs;
    HG_TRY(dev_alloc(c, &s.anno_pool, cap, "annotation pool"));
    HG_TRY(dev_alloc(c, &s.hinge_keep, cap, "hinge flags"));
    s.anno_cap = cap;
    return HG_OK;
}

int hg_set_overlaps(hg_ctx* c, int64_t novl, const int32_t* aread, const int32_t* bread,
                    const int32_t* abpos, const int32_t* aepos, const int32_t* bbpos,
                    const int32_t* bepos, const int32_t* diffs, const int32_t* flags,
                    const int64_t* trace_off, const uint8_t* trace, int32_t tbytes, int32_t where,
                    int32_t a_lo, int32_t a_hi) {
    (void)diffs;  // part of the record layout, read by no stage
    if (!c || c->n_read <= 0) return set_err(c, HG_ERR_ARG, "hg_set_overlaps: call hg_set_reads first");
    if (novl < 0 || (novl > 0 && (!aread || !bread || !abpos || !aepos || !bbpos || !bepos || !flags)))
        return set_err(c, HG_ERR_ARG, "hg_set_overlaps: null column");
    if (a_lo < 0 || a_hi > c->n_read || a_lo >= a_hi)
        return set_err(c, HG_ERR_ARG, "hg_set_overlaps: bad read range");
    if (novl == 0) return set_err(c, HG_ERR_NO_ALIGNMENTS, "No alignments!");
    if (tbytes != 1 && tbytes != 2) return set_err(c, HG_ERR_ARG, "tbytes must be 1 or 2");
    cudaSetDevice(c->device);
    cudaStream_t st = c->stream;
    // host-path buffers are kept across calls when the new batch fits
    const bool reuse = where == HG_MEM_HOST && !c->adopted && c->d_aread != nullptr &&
                       novl <= c->cap_novl;
    if (!reuse) free_overlaps(c);
    c->novl = novl;
    c->a_lo = a_lo;
    c->a_hi = a_hi;
    c->tbytes = tbytes;
    c->has_trace = trace_off != nullptr && trace != nullptr;
    int first_a = 0, last_a = 0;
    if (where == HG_MEM_DEVICE) {
        c->adopted = true;
        const int32_t* cols[7] = {aread, bread, abpos, aepos, bbpos, bepos, flags};
        for (int i = 0; i < 7; i++)
            if (((uintptr_t)cols[i]) & 15)
                return set_err(c, HG_ERR_ARG, "device columns must be 16-byte aligned");
        c->d_aread = (int32_t*)aread; c->d_bread = (int32_t*)bread; c->d_abpos = (int32_t*)abpos;
        c->d_aepos = (int32_t*)aepos; c->d_bbpos = (int32_t*)bbpos; c->d_bepos = (int32_t*)bepos;
        c->d_flags = (int32_t*)flags;
        c->d_trace_off = (int64_t*)trace_off;
        c->d_trace = (uint8_t*)trace;
        HG_TRY(cuda_check(c, cudaMemcpyAsync(&first_a, aread, 4, cudaMemcpyDeviceToHost, st), "D2H"));
        HG_TRY(cuda_check(c, cudaMemcpyAsync(&last_a, aread + (novl - 1), 4, cudaMemcpyDeviceToHost, st), "D2H"));
        HG_TRY(cuda_check(c, cudaStreamSynchronize(st), "D2H"));
    } else {
        c->adopted = false;
        const size_t nb = sizeof(int32_t) * (size_t)novl;
        int32_t** dst[7] = {&c->d_aread, &c->d_bread, &c->d_abpos, &c->d_aepos,
                            &c->d_bbpos, &c->d_bepos, &c->d_flags};
        const int32_t* src[7] = {aread, bread, abpos, aepos, bbpos, bepos, flags};
        for (int i = 0; i < 7; i++) {
            if (!reuse) HG_TRY(dev_alloc(c, dst[i], (size_t)novl + 8, "overlap column"));
            HG_TRY(cuda_check(c, cudaMemcpyAsync(*dst[i], src[i], nb, cudaMemcpyHostToDevice, st), "overlap H2D"));
        }
        if (!reuse) c->cap_novl = novl;
        if (c->has_trace) {
            const int64_t tb = trace_off[novl];
            if (!reuse || !c->d_trace_off) HG_TRY(dev_alloc(c, &c->d_trace_off, (size_t)c->cap_novl + 1, "trace offsets"));
            if (!c->d_trace || tb > c->cap_trace) {
                HG_TRY(dev_alloc(c, &c->d_trace, (size_t)tb + 16, "trace"));
                c->cap_trace = tb;
            }
            HG_TRY(cuda_check(c, cudaMemcpyAsync(c->d_trace_off, trace_off, 8 * ((size_t)novl + 1), cudaMemcpyHostToDevice, st), "trace_off H2D"));
            if (tb > 0) HG_TRY(cuda_check(c, cudaMemcpyAsync(c->d_trace, trace, (size_t)tb, cudaMemcpyHostToDevice, st), "trace H2D"));
        }
        first_a = aread[0];
        last_a = aread[novl - 1];
    }
    c->r_begin = first_a;
    c->r_end = last_a;
    // a global range set earlier (hg_set_global_range) stays in force while the records fit in it
    if (c->g_begin >= 0 && first_a >= c->g_begin && last_a <= c->g_end) {
        c->r_begin = c->g_begin;
        c->r_end = c->g_end;
    } else {
        c->g_begin = c->g_end = -1;
    }
    // CSR + validation + deepest pile-up
    cudaMemsetAsync(c->d_err, 0, sizeof(int) * 4, st);
    launch_csr_validate(c->rec_view(), c->read_view(), c->d_read_off, c->fs.self_cnt, c->d_err, st);
    launch_max_pileup(c->d_read_off, c->n_read, c->d_err + 1, st);
    int h[2] = {0, 0};
    HG_TRY(cuda_check(c, cudaMemcpyAsync(h, c->d_err, 8, cudaMemcpyDeviceToHost, st), "D2H"));
    HG_TRY(cuda_check(c, cudaStreamSynchronize(st), "csr build"));
    if (h[0])
        return set_err(c, HG_ERR_INPUT,
                       "overlap records are not sorted by A-read or violate 0 <= abpos < aepos <= "
                       "rlen[aread], 0 <= bbpos <= bepos <= rlen[bread]");
    c->max_pileup = h[1];

    // scratch that depends on the shape of the data
    FilterScratch& s = c->fs;
    if (s.anno_cap == 0) HG_TRY(alloc_anno_pool(c, c->anno_pool_hint > 0 ? c->anno_pool_hint : 2 * (a_hi - a_lo) + (1 << 16)));
    // K4: one slot of 56 B per pile-up record and warp
    s.hinge_cap = (std::max(c->max_pileup, 32) + 3) & ~3;  // keeps every slot 16-byte aligned
    {
        const size_t slot = (size_t)s.hinge_cap * 56;
        size_t warps = (size_t)c->num_sms * 48;  // the kernel is latency bound: many warps, few reads each
        const size_t budget = (size_t)768 << 20;
        if (warps * slot > budget) warps = std::max<size_t>(32, budget / slot);
        warps = std::max<size_t>(32, warps & ~(size_t)3);  // >= 8 CTA slots: the exact-order tiers share them
        s.hinge_warps = (int)warps;
        if (warps * slot > c->cap_hinge_scratch) {
            HG_TRY(dev_alloc(c, &s.hinge_scratch, warps * slot, "hinge scratch"));
            c->cap_hinge_scratch = warps * slot;
        }
    }
    c->filter_done = false;
    c->shape_version++;
    return HG_OK;
}

static int configure_filter(hg_ctx* c, const hg_filter_params* p) {
    if (p->reso != kReso || p->coverage_fraction == 0)
        return set_err(c, HG_ERR_ARG, "reso is fixed at 40 (filter.cpp:386) and coverage_frac_repeat_annotation must be != 0");
    if (c->filter_params_set && c->configured_shape == c->shape_version &&
        memcmp(&c->fp, p, sizeof(*p)) == 0)
        return HG_OK;  // same parameters, same data shape: launch configuration is still valid
    c->fp = *p;
    c->filter_params_set = true;
    c->configured_shape = c->shape_version;
    FilterScratch& s = c->fs;
    auto bins = [&](int rlen) { return bins_needed(rlen, p->cut_off); };
    {
        // flat K2: pack the reads that have records, [r_begin, r_end] within the owned range, into
        // batches of kFlatBins histogram words; the plan only depends on read lengths and cut_off
        const int lo = std::max(c->a_lo, c->r_begin), hi = std::min(c->a_hi, c->r_end + 1);  // r_begin / r_end: global when set
        if (c->plan_reads_version != c->reads_version || c->plan_cut_off != p->cut_off ||
            c->plan_lo != lo || c->plan_hi != hi || s.flat_capped != (s.flat_kernel == 3) ||
            (s.flat_capped && c->plan_shape != c->shape_version)) {
            // the third form of K1 wants batches that are also bounded by record volume: the CSR comes
            // back for that plan (which then depends on the records, not only on the read lengths)
            const bool capped = s.flat_kernel == 3;
            std::vector<int64_t> h_off;
            if (capped) {
                h_off.resize((size_t)c->n_read + 1);
                HG_TRY(cuda_check(c, cudaMemcpyAsync(h_off.data(), c->d_read_off, sizeof(int64_t) * h_off.size(), cudaMemcpyDeviceToHost, c->stream), "D2H"));
                HG_TRY(cuda_check(c, cudaStreamSynchronize(c->stream), "D2H"));
            }
            FlatPlan plan;
            flat_plan(c->h_rlen.data(), capped ? h_off.data() : nullptr, lo, std::max(lo, hi), c->n_read, p->cut_off, capped, &plan);
            s.flat_capped = capped;
            const std::vector<int2>& batch = plan.batch;
            s.flat_nbatch = hi > lo ? (int)batch.size() - 1 : 0;
            s.flat_lo = lo;
            s.flat_hi = std::max(lo, hi);
            s.flat_bins_total = 0;
            for (const int2& bt : batch) s.flat_bins_total += bt.y;
            const size_t nb1 = (size_t)std::max(s.flat_nbatch, 1);
            HG_TRY(dev_alloc(c, &s.flat_batch, batch.size(), "flat batches"));
            HG_TRY(dev_alloc(c, &s.flat_rbase, plan.rbase.size() + 8, "flat read offsets"));  // + slack, as above
            HG_TRY(dev_alloc(c, &s.flat_rbatch, plan.rbatch.size(), "flat read batches"));
            HG_TRY(dev_alloc(c, &s.flat_desc, plan.desc.size(), "flat batch descriptors"));
            HG_TRY(dev_alloc(c, &s.flat_prof, nb1 * kFlatBins, "coverage profiles"));
            HG_TRY(dev_alloc(c, &s.flat_zmap, nb1 * (kFlatBins / 16), "coverage bit maps"));
            HG_TRY(dev_alloc(c, &s.flat_cmap, nb1 * (kFlatBins / 16), "coverage bit maps"));
            HG_TRY(cuda_check(c, cudaMemcpyAsync(s.flat_batch, batch.data(), sizeof(int2) * batch.size(), cudaMemcpyHostToDevice, c->stream), "H2D"));
            HG_TRY(cuda_check(c, cudaMemcpyAsync(s.flat_rbase, plan.rbase.data(), sizeof(int) * plan.rbase.size(), cudaMemcpyHostToDevice, c->stream), "H2D"));
            HG_TRY(cuda_check(c, cudaMemcpyAsync(s.flat_rbatch, plan.rbatch.data(), sizeof(int) * plan.rbatch.size(), cudaMemcpyHostToDevice, c->stream), "H2D"));
            HG_TRY(cuda_check(c, cudaMemcpyAsync(s.flat_desc, plan.desc.data(), sizeof(int4) * plan.desc.size(), cudaMemcpyHostToDevice, c->stream), "H2D"));
            // per-read results of the reads outside the planned range: "no pile-up"
            cudaMemsetAsync(s.cov_maxbin, 0xff, sizeof(int) * c->n_read, c->stream);
            cudaMemsetAsync(s.mean_cov, 0xff, sizeof(int) * c->n_read, c->stream);
            cudaMemsetAsync(s.rflags, 0, c->n_read, c->stream);
            if (!c->keep_masks) cudaMemsetAsync(s.mask, 0, sizeof(int2) * c->n_read, c->stream);
            cudaMemsetAsync(s.cmask, 0, sizeof(int2) * c->n_read, c->stream);
            cudaMemsetAsync(s.anno_ref, 0, sizeof(int2) * c->n_read, c->stream);
            HG_TRY(cuda_check(c, cudaStreamSynchronize(c->stream), "flat plan"));  // the vectors go away
            c->plan_reads_version = c->reads_version;
            c->plan_shape = c->shape_version;
            c->plan_cut_off = p->cut_off;
            c->plan_lo = lo;
            c->plan_hi = hi;
        }
    }
    s.flat_bins_per_record = (float)(s.flat_bins_total / (double)std::max<int64_t>(c->novl, 1));
    // the generic path is always armed: very long reads, very deep pile-ups, and reads with more
    // raw annotations than the fast path keeps are rerouted to it at run time
    s.big_slot_words = (bins(c->max_rlen) + 31) & ~31;
    s.big_warps = 64;
    HG_TRY(dev_alloc(c, &s.big_scratch, (size_t)s.big_warps * s.big_slot_words, "big-read scratch"));
    return HG_OK;
}

int hg_filter_phase1(hg_ctx* c, const hg_filter_params* p) {
    if (!c || !p || c->novl <= 0) return set_err(c, HG_ERR_ARG, "hg_filter: no overlaps loaded");
    cudaSetDevice(c->device);
    HG_TRY(configure_filter(c, p));
    if (c->peer.world > 1 && !c->peer_connected)
        return set_err(c, HG_ERR_ARG, "hg_peer_export without hg_peer_connect");
    cudaEventRecord(c->ev0, c->stream);
    c->mark(0);
    launch_profile(c->rec_view(), c->read_view(), c->fp, c->r_begin, c->r_end, c->fs, c->stream);
    if (c->peer.world > 1) {
        // sharded, peer exchange: the rank's part of the coverage histogram goes straight into every
        // rank's exchange block
        c->peer.epoch++;
        launch_median(c->read_view(), c->fp, c->fs, 1, c->stream);
        launch_peer_hist_push(c->fs, c->peer, c->stream);
    } else if (c->ext_med_hist) {
        // sharded, NCCL exchange: the rank's part, summed across ranks by the caller
        launch_median(c->read_view(), c->fp, c->fs, 1, c->stream);
    }
    c->mark(1);
    return cuda_check(c, cudaGetLastError(), "filter phase 1");
}

int hg_filter_phase2(hg_ctx* c) {
    if (!c || !c->filter_params_set) return set_err(c, HG_ERR_ARG, "phase2 before phase1");
    cudaSetDevice(c->device);
    cudaStream_t st = c->stream;
    FilterScratch& s = c->fs;
    int* cov0 = nullptr;
    if (c->keep_cov) {
        // profile lengths are known after phase 1: lay the dump out as a CSR
        std::vector<int> maxbin(c->n_read);
        HG_TRY(cuda_check(c, cudaMemcpyAsync(maxbin.data(), s.cov_maxbin, sizeof(int) * c->n_read, cudaMemcpyDeviceToHost, st), "D2H"));
        HG_TRY(cuda_check(c, cudaStreamSynchronize(st), "D2H"));
        c->h_cov0_off.assign((size_t)c->n_read + 1, 0);
        for (int i = 0; i < c->n_read; i++) {
            const bool in = i >= c->a_lo && i < c->a_hi && i >= c->r_begin && i <= c->r_end;
            c->h_cov0_off[i + 1] = c->h_cov0_off[i] + (in ? maxbin[i] + 1 : 0);
        }
        HG_TRY(dev_alloc(c, &c->d_cov0, (size_t)c->h_cov0_off[c->n_read], "coverage dump"));
        HG_TRY(dev_alloc(c, &c->d_cov0_off, (size_t)c->n_read + 1, "coverage offsets"));
        HG_TRY(cuda_check(c, cudaMemcpyAsync(c->d_cov0_off, c->h_cov0_off.data(), 8 * ((size_t)c->n_read + 1), cudaMemcpyHostToDevice, st), "H2D"));
        cov0 = c->d_cov0;
    }
    c->mark(2);
    if (c->peer.world > 1)
        launch_peer_median_pick(c->fp, s, c->peer, st);  // waits for the parts of all ranks
    else
        launch_median(c->read_view(), c->fp, s, c->ext_med_hist ? 2 : 0, st);
    c->mark(3);
    launch_mask_anno(c->rec_view(), c->read_view(), c->fp, c->r_begin, c->r_end, s, cov0,
                     c->d_cov0_off, c->peer, st);
    if (c->peer.world > 1) launch_peer_signal_masks(s, c->peer, st);
    c->mark(4);
    return cuda_check(c, cudaGetLastError(), "filter phase 2");
}

// The launches of phase 3; nothing waits for them here.
static int filter_phase3_launch(hg_ctx* c) {
    if (!c || !c->filter_params_set) return set_err(c, HG_ERR_ARG, "phase3 before phase1");
    cudaSetDevice(c->device);
    cudaStream_t st = c->stream;
    c->mark(5);
    launch_hinge_call(c->rec_view(), c->read_view(), c->fp, c->fs, c->peer, st);
    c->mark(6);
    cudaEventRecord(c->ev1, st);
    return cuda_check(c, cudaGetLastError(), "filter phase 3");
}

int hg_filter_enqueue(hg_ctx* c, const hg_filter_params* p) {
    HG_TRY(hg_filter_phase1(c, p));
    HG_TRY(hg_filter_phase2(c));
    return filter_phase3_launch(c);
}

int hg_filter_phase3(hg_ctx* c, hg_filter_summary* out) {
    HG_TRY(filter_phase3_launch(c));
    return hg_filter_finish(c, out);
}

// Waits for the stage (the last one enqueued) and reads its counters back.
int hg_filter_finish(hg_ctx* c, hg_filter_summary* out) {
    if (!c || !c->filter_params_set) return set_err(c, HG_ERR_ARG, "hg_filter_finish before a filter run");
    cudaSetDevice(c->device);
    cudaStream_t st = c->stream;
    FilterScratch& s = c->fs;
    int cnt[16], scal[8];
    HG_TRY(cuda_check(c, cudaMemcpyAsync(cnt, s.counters, sizeof cnt, cudaMemcpyDeviceToHost, st), "D2H"));
    HG_TRY(cuda_check(c, cudaMemcpyAsync(scal, s.scal, sizeof scal, cudaMemcpyDeviceToHost, st), "D2H"));
    HG_TRY(cuda_check(c, cudaStreamSynchronize(st), "filter"));
    if (cnt[15])
        return set_err(c, HG_ERR_CUDA, "peer exchange timed out: a rank did not reach the same hg_filter call");
    if (scal[5])
        return set_err(c, HG_ERR_INPUT, "median coverage >= 4095: not supported with a shared coverage histogram "
                                        "(HG_BUF_MEDIAN_HIST / peer exchange); all-gather HG_BUF_MEAN_COV instead");
    if (cnt[2]) {
        // annotation pool overflow (here, or on another rank of a peer-connected run): grow it if it
        // was ours and tell the caller to rerun the three phases
        if (cnt[0] > s.anno_cap) HG_TRY(alloc_anno_pool(c, std::max(cnt[0] + (1 << 16), s.anno_cap * 2)));
        return HG_RETRY_POOL;
    }
    c->filter_done = true;
    if (out) {
        out->r_begin = c->r_begin;
        out->r_end = c->r_end;
        out->cov_est = scal[0];
        out->min_cov = scal[1];
        out->n_annotations = cnt[0];
        out->n_hinges = -1;
        out->n_exact_order = cnt[4];
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        out->ms_device = ms;
    }
    return HG_OK;
}

int hg_filter(hg_ctx* c, const hg_filter_params* p, hg_filter_summary* out) {
    for (int attempt = 0; attempt < 8; attempt++) {
        HG_TRY(hg_filter_phase1(c, p));
        HG_TRY(hg_filter_phase2(c));
        const int rc = hg_filter_phase3(c, out);
        if (rc != HG_RETRY_POOL) return rc;  // else: the pool has been grown, run again
    }
    return set_err(c, HG_ERR_NOMEM, "annotation pool kept overflowing");
}

// ms[0..3] = coverage profiles (K1), median, mask + annotation (K2), hinge calls (K4)
int hg_filter_kernel_times(hg_ctx* c, float* ms, int n) {
    if (!c || !c->profile || !c->filter_done) return set_err(c, HG_ERR_ARG, "profiling is off");
    const int pairs[4][2] = {{0, 1}, {2, 3}, {3, 4}, {5, 6}};
    for (int i = 0; i < n && i < 4; i++) cudaEventElapsedTime(&ms[i], c->marks[pairs[i][0]], c->marks[pairs[i][1]]);
    return HG_OK;
}

int64_t hg_launch_count(void) { return hg::g_launches; }

// Test hooks for the order-exact sort: `count` arrays laid end to end (off has count+1 entries),
// each element = (key, idx).  The device version runs warp_sort_exact, the host version the real
// std::sort of this toolchain; tests/test_gpu_order_exact.py compares them element for element.
int hg_debug_warp_sort(hg_ctx* c, int32_t* key_idx, const int32_t* off, int32_t count, int32_t descending) {
    if (!c || !key_idx || !off || count <= 0) return HG_ERR_ARG;
    cudaSetDevice(c->device);
    const size_t total = (size_t)off[count];
    int2* d = nullptr; int2* tmp = nullptr; int *g = nullptr, *l = nullptr, *doff = nullptr;
    HG_TRY(dev_alloc(c, &d, total, "sort data")); HG_TRY(dev_alloc(c, &tmp, total, "sort tmp"));
    HG_TRY(dev_alloc(c, &g, total, "sort g")); HG_TRY(dev_alloc(c, &l, total, "sort l"));
    HG_TRY(dev_alloc(c, &doff, (size_t)count + 1, "sort off"));
    cudaMemcpy(d, key_idx, 8 * total, cudaMemcpyHostToDevice);
    cudaMemcpy(doff, off, 4 * ((size_t)count + 1), cudaMemcpyHostToDevice);
    launch_debug_warp_sort(d, doff, count, descending, g, l, tmp, c->stream);
    int rc = cuda_check(c, cudaStreamSynchronize(c->stream), "warp sort");
    if (rc == HG_OK) cudaMemcpy(key_idx, d, 8 * total, cudaMemcpyDeviceToHost);
    cudaFree(d); cudaFree(tmp); cudaFree(g); cudaFree(l); cudaFree(doff);
    return rc;
}

// Test hook (no GPU needed): the host-side batch plan of the flat filter kernels.
// batch_out gets (first read, histogram words in use) pairs, closed by (hi, 0); returns the number
// of batches, or -1 when capacity (in pairs) is too small.
int hg_debug_flat_plan(const int32_t* rlen, int32_t n_read, int32_t lo, int32_t hi, int32_t cut_off,
                       int32_t* batch_out, int32_t capacity, int32_t* rbase_out) {
    FlatPlan plan;
    flat_plan(rlen, nullptr, lo, hi, n_read, cut_off, false, &plan);
    const std::vector<int2>& batch = plan.batch;
    const std::vector<int>& rbase = plan.rbase;
    if ((int)batch.size() > capacity) return -1;
    for (size_t i = 0; i < batch.size(); i++) {
        batch_out[2 * i] = batch[i].x;
        batch_out[2 * i + 1] = batch[i].y;
    }
    memcpy(rbase_out, rbase.data(), sizeof(int) * (size_t)n_read);
    return (int)batch.size() - 1;
}

// The same with the CSR (batches also bounded by record volume): nrec_out gets the records per batch.
int hg_debug_flat_plan2(const int32_t* rlen, const int64_t* read_off, int32_t n_read, int32_t lo, int32_t hi,
                        int32_t cut_off, int32_t* batch_out, int32_t capacity, int32_t* rbase_out,
                        int32_t* rbatch_out, int32_t* nrec_out) {
    FlatPlan plan;
    flat_plan(rlen, read_off, lo, hi, n_read, cut_off, true, &plan);
    if ((int)plan.batch.size() > capacity) return -1;
    for (size_t i = 0; i < plan.batch.size(); i++) {
        batch_out[2 * i] = plan.batch[i].x;
        batch_out[2 * i + 1] = plan.batch[i].y;
    }
    memcpy(rbase_out, plan.rbase.data(), sizeof(int) * (size_t)n_read);
    memcpy(rbatch_out, plan.rbatch.data(), sizeof(int) * (size_t)n_read);
    for (size_t i = 0; i + 1 < plan.batch.size(); i++) nrec_out[i] = plan.desc[2 * i + 1].z;
    return (int)plan.batch.size() - 1;
}

int hg_debug_std_sort(int32_t* key_idx, const int32_t* off, int32_t count, int32_t descending) {
    struct E { int key, idx; };
    for (int w = 0; w < count; w++) {
        E* b = reinterpret_cast<E*>(key_idx) + off[w];
        E* e = reinterpret_cast<E*>(key_idx) + off[w + 1];
        if (descending)
            std::sort(b, e, [](const E& x, const E& y) { return x.key > y.key; });
        else
            std::sort(b, e, [](const E& x, const E& y) { return x.key < y.key; });
    }
    return HG_OK;
}

// Debug aid (HG_OPT_PROFILE set after hg_set_reads): per hinge-call work item
// (read, cycles, largest support, pile-up size if the order-exact path ran).
int hg_debug_item_log(hg_ctx* c, int32_t* out4, int64_t capacity, int64_t* n_items) {
    if (!c || !c->fs.item_log || !n_items) return HG_ERR_ARG;
    int cnt[8];
    cudaMemcpy(cnt, c->fs.counters, sizeof cnt, cudaMemcpyDeviceToHost);
    *n_items = cnt[1];
    if (out4 && capacity >= cnt[1])
        cudaMemcpy(out4, c->fs.item_log, sizeof(int4) * (size_t)cnt[1], cudaMemcpyDeviceToHost);
    return HG_OK;
}

int hg_bind_buffer(hg_ctx* c, int which, void* dptr, int64_t bytes) {
    if (!c || !dptr || c->n_read <= 0) return HG_ERR_ARG;
    if (which == HG_BUF_MEAN_COV && bytes >= 4ll * c->n_read) {
        if (!c->ext_mean_cov) cudaFree(c->fs.mean_cov);
        c->fs.mean_cov = (int*)dptr;
        c->ext_mean_cov = true;
        c->plan_reads_version = -1;  // the next run clears the entries outside its read range
        c->filter_params_set = false;
        return HG_OK;
    }
    if (which == HG_BUF_MEDIAN_HIST && bytes >= 4ll * 4098) {
        if (!c->ext_med_hist) cudaFree(c->fs.med_hist);
        c->fs.med_hist = (unsigned int*)dptr;
        c->ext_med_hist = true;
        return cuda_check(c, cudaMemset(dptr, 0, 4 * 4098), "median histogram");
    }
    if (which == HG_BUF_MASK_PACKED && bytes >= 4ll * c->n_read) {
        // every mask bound is a multiple of gcd(40, tspace): coverage bins and QV tiles
        int g = kReso, t = c->tspace > 0 ? c->tspace : kReso;
        while (t) {
            const int r = g % t;
            g = t;
            t = r;
        }
        if (c->max_rlen / g >= 65536)
            return set_err(c, HG_ERR_ARG, "HG_BUF_MASK_PACKED: reads too long for 16-bit mask bounds; bind HG_BUF_MASK");
        c->fs.mask_pk = (uint32_t*)dptr;
        c->fs.mask_g = g;
        return cuda_check(c, cudaMemset(dptr, 0, 4ull * c->n_read), "packed masks");
    }
    if (which == HG_BUF_MASK && bytes >= 8ll * c->n_read) {
        if (!c->ext_mask) cudaFree(c->fs.mask);
        c->fs.mask = (int2*)dptr;
        c->ext_mask = true;
        c->plan_reads_version = -1;
        c->filter_params_set = false;
        return HG_OK;
    }
    return set_err(c, HG_ERR_ARG, "hg_bind_buffer: unknown buffer or too small");
}

int hg_device_buffer(hg_ctx* c, int which, void** dptr, int64_t* bytes) {
    if (!c || !dptr || !bytes || c->n_read <= 0) return HG_ERR_ARG;
    switch (which) {
        case HG_BUF_MEAN_COV: *dptr = c->fs.mean_cov; *bytes = 4ll * c->n_read; return HG_OK;
        case HG_BUF_MASK: *dptr = c->fs.mask; *bytes = 8ll * c->n_read; return HG_OK;
        case HG_BUF_READ_FLAGS: *dptr = c->fs.rflags; *bytes = c->n_read; return HG_OK;
        default: return set_err(c, HG_ERR_ARG, "unknown buffer");
    }
}

int hg_set_global_range(hg_ctx* c, int32_t first_aread, int32_t last_aread) {
    if (!c || c->novl <= 0) return set_err(c, HG_ERR_ARG, "hg_set_global_range: call hg_set_overlaps first");
    if (first_aread < 0 || last_aread >= c->n_read || first_aread > last_aread)
        return set_err(c, HG_ERR_ARG, "hg_set_global_range: range does not contain this context's records");
    c->g_begin = first_aread;
    c->g_end = last_aread;
    c->r_begin = first_aread;
    c->r_end = last_aread;
    c->shape_version++;
    return HG_OK;
}

// ---- peer exchange ---------------------------------------------------------------------

static int peer_alloc_block(hg_ctx* c, int rank, int world) {
    if (c->n_read <= 0) return set_err(c, HG_ERR_ARG, "hg_peer_export: call hg_set_reads first");
    if (world < 2 || world > kMaxPeers || rank < 0 || rank >= world)
        return set_err(c, HG_ERR_ARG, "hg_peer_export: 2 <= world <= 16, 0 <= rank < world");
    if (c->ext_med_hist || (c->fs.mask_pk && !c->peer_block))
        return set_err(c, HG_ERR_ARG, "hg_peer_export: context already bound to NCCL-exchanged buffers");
    int g = kReso, t = c->tspace > 0 ? c->tspace : kReso;
    while (t) {
        const int r = g % t;
        g = t;
        t = r;
    }
    if (c->max_rlen / g >= 65536)
        return set_err(c, HG_ERR_ARG, "peer exchange: reads too long for 16-bit mask bounds; use the NCCL exchange with HG_BUF_MASK");
    cudaSetDevice(c->device);
    peer_teardown(c);
    c->peer_block_bytes = kPeerFlagBytes + kPeerHistBytes + sizeof(uint32_t) * (size_t)c->n_read;
    HG_TRY(cuda_check(c, cudaMalloc((void**)&c->peer_block, c->peer_block_bytes), "exchange block"));
    HG_TRY(cuda_check(c, cudaMemset(c->peer_block, 0, c->peer_block_bytes), "exchange block"));
    HG_TRY(cuda_check(c, cudaDeviceSynchronize(), "exchange block"));
    c->peer.rank = rank;
    c->peer.world = world;
    c->peer.epoch = 0;
    c->peer.base[rank] = c->peer_block;
    c->fs.mask_pk = c->peer.mask_pk(rank);
    c->fs.mask_g = g;
    c->peer_connected = false;
    return HG_OK;
}

int hg_peer_export(hg_ctx* c, int32_t rank, int32_t world, void* handle_out) {
    if (!c || !handle_out) return HG_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) <= HG_PEER_HANDLE_BYTES, "handle size");
    HG_TRY(peer_alloc_block(c, rank, world));
    cudaIpcMemHandle_t h;
    HG_TRY(cuda_check(c, cudaIpcGetMemHandle(&h, c->peer_block), "cudaIpcGetMemHandle"));
    memset(handle_out, 0, HG_PEER_HANDLE_BYTES);
    memcpy(handle_out, &h, sizeof h);
    return HG_OK;
}

int hg_peer_connect(hg_ctx* c, const void* handles) {
    if (!c || !handles || c->peer.world < 2 || !c->peer_block)
        return set_err(c, HG_ERR_ARG, "hg_peer_connect: call hg_peer_export first");
    cudaSetDevice(c->device);
    for (int r = 0; r < c->peer.world; r++) {
        if (r == c->peer.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const uint8_t*)handles + (size_t)r * HG_PEER_HANDLE_BYTES, sizeof h);
        void* p = nullptr;
        HG_TRY(cuda_check(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle"));
        c->peer.base[r] = (uint8_t*)p;
        c->peer_ipc[r] = true;
    }
    c->peer_connected = true;
    return HG_OK;
}

int hg_peer_connect_local(hg_ctx** ctxs, int32_t world) {
    if (!ctxs || world < 2 || world > kMaxPeers) return HG_ERR_ARG;
    for (int r = 0; r < world; r++) {
        if (!ctxs[r]) return HG_ERR_ARG;
        HG_TRY(peer_alloc_block(ctxs[r], r, world));
    }
    for (int r = 0; r < world; r++) {
        hg_ctx* c = ctxs[r];
        cudaSetDevice(c->device);
        for (int q = 0; q < world; q++) {
            if (q == r) continue;
            if (ctxs[q]->device != c->device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, c->device, ctxs[q]->device);
                if (!can) return set_err(c, HG_ERR_CUDA, "no peer access between the devices");
                const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[q]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_check(c, e, "cudaDeviceEnablePeerAccess");
                cudaGetLastError();
            }
            c->peer.base[q] = ctxs[q]->peer_block;
        }
        c->peer_connected = true;
    }
    return HG_OK;
}

int hg_peer_masks(hg_ctx* c, int32_t* mask) {
    if (!c || !mask || !c->peer_block || !c->filter_done) return set_err(c, HG_ERR_ARG, "hg_peer_masks: no peer-connected run");
    cudaSetDevice(c->device);
    std::vector<uint32_t> pk((size_t)c->n_read);
    HG_TRY(cuda_check(c, cudaMemcpyAsync(pk.data(), c->fs.mask_pk, 4ull * c->n_read, cudaMemcpyDeviceToHost, c->stream), "D2H"));
    HG_TRY(cuda_check(c, cudaStreamSynchronize(c->stream), "D2H"));
    for (int i = 0; i < c->n_read; i++) {
        mask[2 * i] = (int)(pk[i] & 0xffffu) * c->fs.mask_g;
        mask[2 * i + 1] = (int)(pk[i] >> 16) * c->fs.mask_g;
    }
    return HG_OK;
}

int hg_filter_fetch(hg_ctx* c, int32_t* mask, int32_t* cmask, uint8_t* flags, int64_t* anno_off,
                    int32_t* anno_pos, int32_t* anno_type, uint8_t* hinge_keep) {
    if (!c || !c->filter_done) return set_err(c, HG_ERR_ARG, "hg_filter_fetch: run hg_filter first");
    cudaSetDevice(c->device);
    cudaStream_t st = c->stream;
    FilterScratch& s = c->fs;
    const int n = c->n_read;
    if (mask) HG_TRY(cuda_check(c, cudaMemcpyAsync(mask, s.mask, 8ull * n, cudaMemcpyDeviceToHost, st), "D2H"));
    if (cmask) HG_TRY(cuda_check(c, cudaMemcpyAsync(cmask, s.cmask, 8ull * n, cudaMemcpyDeviceToHost, st), "D2H"));
    if (flags) HG_TRY(cuda_check(c, cudaMemcpyAsync(flags, s.rflags, n, cudaMemcpyDeviceToHost, st), "D2H"));
    if (anno_off) {
        int used = 0;
        std::vector<int2> ref(n);
        HG_TRY(cuda_check(c, cudaMemcpyAsync(&used, s.counters, 4, cudaMemcpyDeviceToHost, st), "D2H"));
        HG_TRY(cuda_check(c, cudaMemcpyAsync(ref.data(), s.anno_ref, 8ull * n, cudaMemcpyDeviceToHost, st), "D2H"));
        HG_TRY(cuda_check(c, cudaStreamSynchronize(st), "D2H"));
        std::vector<int2> pool((size_t)std::max(used, 1));
        std::vector<uint8_t> keep((size_t)std::max(used, 1));
        if (used > 0) {
            HG_TRY(cuda_check(c, cudaMemcpyAsync(pool.data(), s.anno_pool, 8ull * used, cudaMemcpyDeviceToHost, st), "D2H"));
            HG_TRY(cuda_check(c, cudaMemcpyAsync(keep.data(), s.hinge_keep, used, cudaMemcpyDeviceToHost, st), "D2H"));
            HG_TRY(cuda_check(c, cudaStreamSynchronize(st), "D2H"));
        }
        // the pool is filled in completion order; hand it back in read order
        int64_t o = 0;
        for (int i = 0; i < n; i++) {
            anno_off[i] = o;
            for (int k = 0; k < ref[i].y; k++, o++) {
                if (anno_pos) anno_pos[o] = pool[ref[i].x + k].x;
                if (anno_type) anno_type[o] = pool[ref[i].x + k].y;
                if (hinge_keep) hinge_keep[o] = keep[ref[i].x + k];
            }
        }
        anno_off[n] = o;
    }
    HG_TRY(cuda_check(c, cudaStreamSynchronize(st), "hg_filter_fetch"));
    if (flags)  // internal bits (hinge pre-test) stay inside the library
        for (int i = 0; i < n; i++) flags[i] &= (kFlagCov | kFlagSelf);
    return HG_OK;
}

int hg_filter_coverage(hg_ctx* c, int64_t* cov_off, int32_t* cov, int64_t* n_bins) {
    if (!c || !c->filter_done || !c->keep_cov)
        return set_err(c, HG_ERR_ARG, "hg_filter_coverage: set HG_OPT_KEEP_COVERAGE before hg_filter");
    const int64_t total = c->h_cov0_off[c->n_read];
    if (n_bins) *n_bins = total;
    if (cov_off) memcpy(cov_off, c->h_cov0_off.data(), 8 * ((size_t)c->n_read + 1));
    if (cov && total > 0) {
        cudaSetDevice(c->device);
        HG_TRY(cuda_check(c, cudaMemcpyAsync(cov, c->d_cov0, 4ull * total, cudaMemcpyDeviceToHost, c->stream), "D2H"));
        HG_TRY(cuda_check(c, cudaStreamSynchronize(c->stream), "D2H"));
    }
    return HG_OK;
}

}  // extern "C"
