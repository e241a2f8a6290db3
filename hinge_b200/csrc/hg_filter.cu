// Kernels of the `hinge filter` stage (sm_100a, integer / index work, no tensor
// cores).  Reference behaviour: /root/reference/src/filter/filter.cpp:529-1098.
//
//   K0 csr_validate    .las order + invariants, CSR offsets by A-read, A == B records per read
//   K0 qv_mask         longest good-QV run per read          (filter.cpp:340-369)
//   K1 / K2            coverage profiles, masks, annotation: hg_filter_flat.cu; here only
//                      the generic per-read path for reads those kernels cannot take
//                      (mask_anno_read, k_mask_anno_big)       (filter.cpp:588-865)
//      median_*        median of the per-read means -> MIN_COV (filter.cpp:660-678)
//   K4 hinge_call      one warp per annotated read: selection, rank sort, tie
//                      detection, bridged / unbridged walk     (filter.cpp:867-1066)
//      hinge_exact     one CTA per read whose result depends on the reference's
//                      unstable sort order: introsort restated (hg_order.h)
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "hg_device.cuh"
#include "hg_filter.h"
#include "hg_order.h"

namespace hg {

int64_t g_launches = 0;

// ------------------------------------------------------------------ K0

__global__ void k_csr_validate(RecView rv, ReadView rd, int64_t* read_off, int* self_cnt, int* err) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > rv.novl) return;
    const int prev = k == 0 ? rd.r_lo - 1 : rv.aread[k - 1];
    int cur = k == rv.novl ? rd.r_hi - 1 : rv.aread[k];
    if (k < rv.novl) {
        const int a = rv.aread[k], b = rv.bread[k];
        bool ok = a >= rd.r_lo && a < rd.r_hi && b >= 0 && b < rd.n_read && a >= prev;
        if (ok) {
            const int as = rv.abpos[k], ae = rv.aepos[k], bs = rv.bbpos[k], be = rv.bepos[k];
            ok = as >= 0 && as < ae && ae <= rd.rlen[a] && bs >= 0 && bs <= be && be <= rd.rlen[b];
        }
        if (!ok) {
            atomicExch(err, 1);
            return;
        }
        if (a == b) atomicAdd(&self_cnt[a], 1);  // inactive in every stage (filter.cpp:538-547); rare
        if (rv.abpos[k] / kReso == rv.aepos[k] / kReso) atomicOr(&self_cnt[a], kSelfDegenerate);
    }
    // every read in (prev, cur] starts at record k
    for (int r = max(prev + 1, rd.r_lo); r <= cur; r++) read_off[r] = k;
    if (k == rv.novl) {
        for (int r = 0; r < rd.r_lo; r++) read_off[r] = 0;  // reads of other shards: empty
        for (int r = rd.r_hi; r <= rd.n_read; r++) read_off[r] = rv.novl;
    }
}

// filter.cpp:340-369: a tile is good when its QV < 40 (filter.cpp:311); a bad
// tile or the LAST tile closes the current run; keep the strictly longest run.
__global__ void k_qv_mask(int n_read, const int64_t* __restrict__ qv_off,
                          const uint8_t* __restrict__ qv, int tspace, int2* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_read) return;
    const int64_t o = qv_off[i];
    const int n = (int)(qv_off[i + 1] - o);
    int s = 0, e = 0, best = 0, bs = 0, be = 0;
    for (int j = 0; j < n; j++) {
        const bool good = qv[o + j] < 40;
        if (good && j < n - 1) {
            e++;
        } else {
            if (e - s > best) {
                be = e;
                bs = s;
                best = e - s;
            }
            s = j + 1;
            e = j + 1;
        }
    }
    out[i] = make_int2(bs * tspace, be * tspace);
}

// ------------------------------------------------------------------ median
// (K1, the profile build, and the flat K2 live in hg_filter_flat.cu)

// Median by counting: per-read means are small integers, so a 4096-bin histogram resolves the
// rank exactly (filter.cpp:660 median_id = size / 2); means >= 4095 (coverage in the thousands)
// take the generic three-digit radix select below.  The block that finishes last picks the
// median and clears the histogram for the next run: one launch, no host round trip.
// scal: [0] cov_est  [1] MIN_COV  [2] radix prefix  [3] radix rank  [4] need radix pass
//
// Sharded runs (hg_bind_buffer(HG_BUF_MEDIAN_HIST)): the accumulation covers the rank's own reads
// (pick = 0), the histograms are summed across ranks (16 KB instead of the per-read means), and
// k_median_pick finishes on every rank; the radix fallback would need all the means, so a
// median in the overflow bin is reported as an error there (scal[5]).
__device__ void median_pick_block(unsigned int* __restrict__ hist, int est_cov, int min_cov,
                                  int* __restrict__ scal, bool allow_radix);

__global__ void __launch_bounds__(256)
k_median_hist(const int* __restrict__ mean_cov, int r_lo, int r_hi, unsigned int* __restrict__ hist /*4096 + 2*/,
              int est_cov, int min_cov, int* __restrict__ scal, int pick) {
    __shared__ unsigned int sh[4096];
    __shared__ int is_last;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    unsigned int valid = 0;
    for (int i = r_lo + blockIdx.x * blockDim.x + threadIdx.x; i < r_hi; i += gridDim.x * blockDim.x) {
        const int v = mean_cov[i];
        if (v >= 0) {
            atomicAdd(&sh[min(v, 4095)], 1u);
            valid++;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4096; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
    valid = (unsigned int)warp_sum((int)valid);
    if (lane_id() == 0 && valid) atomicAdd(&hist[4096], valid);
    if (!pick) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(&hist[4097], 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    median_pick_block(hist, est_cov, min_cov, scal, true);
}

__global__ void __launch_bounds__(256)
k_median_pick(unsigned int* __restrict__ hist, int est_cov, int min_cov, int* __restrict__ scal) {
    median_pick_block(hist, est_cov, min_cov, scal, false);
}

// One CTA of 256 threads: thread t owns bins [16 t, 16 t + 16); clears the histogram afterwards.
__device__ void median_pick_block(unsigned int* __restrict__ hist, int est_cov, int min_cov,
                                  int* __restrict__ scal, bool allow_radix) {
    __shared__ unsigned int wsum[8];
    const int t = threadIdx.x;
    unsigned int loc[16], tot = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        loc[i] = *(volatile unsigned int*)&hist[16 * t + i];  // global (written by other CTAs) or shared
        tot += loc[i];
    }
    const unsigned int m = *(volatile unsigned int*)&hist[4096];
    const unsigned int incl = warp_incl_scan(tot);
    if (lane_id() == 31) wsum[t >> 5] = incl;
    __syncthreads();
    unsigned int before = incl - tot;
    for (int w = 0; w < (t >> 5); w++) before += wsum[w];
    const unsigned int rank = m / 2;
    if (m == 0) {
        if (t == 0) {
            const int cov_est = est_cov != 0 ? est_cov : 0;  // filter.cpp:671
            scal[0] = cov_est;
            scal[1] = max(min_cov, cov_est / 3);
            scal[2] = 0;
            scal[4] = 0;
            scal[5] = 0;
        }
    } else if (before <= rank && rank < before + tot) {  // exactly one thread
        unsigned int run = before;
        int v = 16 * t;
        for (int i = 0; i < 16; i++, v++) {
            if (run + loc[i] > rank) break;
            run += loc[i];
        }
        const int need = v >= 4095;  // inside the overflow bin: resolve with the radix passes
        scal[2] = 0;
        scal[3] = (int)(rank - run);
        scal[4] = need && allow_radix;
        scal[5] = need && !allow_radix;
        if (!need || !allow_radix) {
            const int cov_est = est_cov != 0 ? est_cov : v;  // filter.cpp:671
            scal[0] = cov_est;
            scal[1] = max(min_cov, cov_est / 3);  // filter.cpp:677-678
        }
    }
    __syncthreads();
    for (int i = t; i < 4098; i += blockDim.x) hist[i] = 0;
}

// Generic fallback, one CTA, three passes (digits 12/10/10 bits of v - 4095, high to low).
// Only does anything when the median lies in the overflow bin.
__global__ void __launch_bounds__(1024)
k_median_radix(const int* __restrict__ mean_cov, int n_read, int est_cov, int min_cov,
               int* __restrict__ scal) {
    if (!scal[4]) return;
    __shared__ unsigned int dig[4096];
    __shared__ unsigned int s_prefix, s_rank;
    if (threadIdx.x == 0) {
        s_prefix = 0;
        s_rank = (unsigned int)scal[3];
    }
    for (int pass = 0; pass < 3; pass++) {
        const int shift = pass == 0 ? 20 : (pass == 1 ? 10 : 0);
        const int bits = pass == 0 ? 12 : 10;
        for (int i = threadIdx.x; i < 4096; i += blockDim.x) dig[i] = 0;
        __syncthreads();
        const unsigned int prefix = s_prefix;
        for (int i = threadIdx.x; i < n_read; i += blockDim.x) {
            const int v = mean_cov[i];
            if (v < 4095) continue;
            const unsigned int u = (unsigned int)(v - 4095);
            const unsigned int hi = pass == 0 ? 0u : (u >> (shift + bits));
            if (hi != prefix) continue;
            atomicAdd(&dig[(u >> shift) & ((1u << bits) - 1u)], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int rank = s_rank, run = 0, d = 0;
            for (; d < (1u << bits); d++) {
                if (run + dig[d] > rank) break;
                run += dig[d];
            }
            s_rank = rank - run;
            s_prefix = (prefix << bits) | d;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int cov_est = (int)(s_prefix + 4095u);
        if (est_cov != 0) cov_est = est_cov;
        scal[0] = cov_est;
        scal[1] = max(min_cov, cov_est / 3);
    }
}

// ------------------------------------------------------------------ peer exchange (sharded runs)

// After k_median_hist (pick = 0) has accumulated the rank's own reads: one CTA stores the rank's
// part of the histogram into row `rank` of EVERY rank's exchange block (plain stores over NVLink,
// 16 KB per peer), clears the local accumulator for the next run, and then raises its arrival
// flag in every block.  Writers fence before the barrier, the flag writers again after it.
__global__ void __launch_bounds__(1024)
k_peer_hist_push(unsigned int* __restrict__ hist, PeerView pv) {
    for (int b = threadIdx.x; b < 4098; b += blockDim.x) {
        const unsigned v = hist[b];
        hist[b] = 0;
        for (int r = 0; r < pv.world; r++) pv.hist(r, pv.rank)[b] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < pv.world) {
        __threadfence_system();
        st_release_sys(pv.flag_hist(threadIdx.x) + pv.rank, pv.epoch);
    }
}

// Waits for the parts of all ranks, sums them and picks the median (one CTA, 256 threads: thread t
// owns bins [16 t, 16 t + 16)).  err[0] = 1 when a peer did not arrive in time.
__global__ void __launch_bounds__(256)
k_peer_median_pick(PeerView pv, int est_cov, int min_cov, int* __restrict__ scal, int* __restrict__ err) {
    __shared__ unsigned int sum[4098];
    if (threadIdx.x < pv.world && !peer_wait(pv.flag_hist(pv.rank) + threadIdx.x, pv.epoch)) atomicExch(err, 1);
    __syncthreads();
    for (int b = threadIdx.x; b < 4098; b += blockDim.x) {
        unsigned v = 0;
        for (int r = 0; r < pv.world; r++) v += __ldcg(pv.hist(pv.rank, r) + b);
        sum[b] = v;
    }
    __syncthreads();
    median_pick_block(sum, est_cov, min_cov, scal, false);
}

// After K2 (both forms): everything this rank stored into the other ranks' mask arrays is
// complete at the kernel boundary; tell every rank, together with the annotation-pool overflow
// bit so that all ranks agree on rerunning the stage (HG_RETRY_POOL).
__global__ void k_peer_signal_masks(PeerView pv, const int* __restrict__ counters) {
    if (threadIdx.x < pv.world) {
        __threadfence_system();
        pv.flag_ovf(threadIdx.x)[pv.rank] = counters[2] ? pv.epoch : pv.epoch - 1;
        __threadfence_system();
        st_release_sys(pv.flag_mask(threadIdx.x) + pv.rank, pv.epoch);
    }
}

// ------------------------------------------------------------------ K2

// One warp owns one read.  `hist` holds `nbz` packed words (shared memory on
// the fast path, global scratch for very long reads / very deep pile-ups).
template <typename W>
__device__ void mask_anno_read(const RecView& rv, const ReadView& rd, const hg_filter_params& P,
                               const int MIN_COV, const int read, W* hist, const int nbz,
                               const MaskAnnoOut& out) {
    typedef Packed<W> PK;
    const int lane = lane_id();
    constexpr int reso = kReso;
    const int64_t o0 = rv.read_off[read], o1 = rv.read_off[read + 1];

    for (int j = lane; j < nbz; j += 32) hist[j] = 0;
    __syncwarp();

    // ---- scatter the four events of every record (A == B records are inactive,
    //      filter.cpp:538-547); profileCoverage, LAInterface.cpp:4298-4320
    int m0 = -1, mc = -1;
    for (int64_t k = o0 + lane; k < o1; k += 32) {
        const int b = __ldg(rv.bread + k);
        if (b == read) continue;
        const int as = __ldg(rv.abpos + k), ae = __ldg(rv.aepos + k);
        const int b_s0 = cov_bin(as, reso), b_e0 = cov_bin(ae, reso);
        const int b_sc = cov_bin(as + P.cut_off, reso), b_ec = cov_bin(ae - P.cut_off, reso);
        PK::add(&hist[b_s0], PK::one_lo());
        PK::add(&hist[b_e0], (W)0 - PK::one_lo());
        PK::add(&hist[b_sc], PK::one_hi());
        PK::add(&hist[b_ec], (W)0 - PK::one_hi());
        m0 = max(m0, b_e0);
        mc = max(mc, max(b_sc, b_ec));
    }
    const int L0 = warp_max(m0) + 1;  // profile lengths (0 for an empty pile-up)
    const int LC = warp_max(mc) + 1;
    __syncwarp();

    // ---- prefix sums in place + longest run of covered bins (filter.cpp:696-728).
    // A bin is "zero" when covC - MIN_COV <= 0; the run between two consecutive
    // zeros p < z scores 40 * (z - p - 2); bin 0 acts as a zero; strict '>' keeps
    // the earliest of the longest runs.
    W carry = 0;
    int last_zero = 0;
    unsigned long long best = 0;  // (gap << 32) | ~z  -> max gap, then smallest z
    for (int base = 0; base < nbz; base += 32) {
        const int j = base + lane;
        W v = j < nbz ? hist[j] : 0;
        v = warp_incl_scan(v) + carry;
        carry = __shfl_sync(0xffffffffu, v, 31);
        if (j < nbz) hist[j] = v;
        const bool zero = j < LC && !(PK::hi(v) > MIN_COV);
        const unsigned zm = __ballot_sync(0xffffffffu, zero);
        if (zero) {
            const unsigned below = zm & ((1u << lane) - 1u);
            const int p = below ? base + 31 - __clz(below) : last_zero;
            const int gap = j - p;
            if (gap >= 3) {
                const unsigned long long cand =
                    ((unsigned long long)gap << 32) | (unsigned)(0x7fffffff - j);
                if (cand > best) best = cand;
            }
        }
        if (zm) last_zero = base + 31 - __clz(zm);
    }
    best = warp_max_u64(best);
    __syncwarp();
    int maxstart = 0, maxend = 0, msc = 0, mec = 0;
    if (best) {
        const int gap = (int)(best >> 32), z = 0x7fffffff - (int)(best & 0xffffffffu), p = z - gap;
        msc = p + 1;
        mec = z - 1;
        maxstart = reso * (p + 1);
        maxend = reso * (z - 1);
    }

    // ---- telomere / coverage-imbalance flag (filter.cpp:731-760)
    uint8_t flags = out.rflags[read] & kFlagSelf;
    if (P.delete_telomere) {
        int limit, div;
        if (mec - msc + 1 > 20) {
            limit = 10;
            div = 10;
        } else {
            limit = (mec - msc) / 2;
            div = limit;
        }
        int sc = 0, ec = 0;
        for (int t = lane; t < limit; t += 32) {
            sc += max(PK::hi(hist[msc + t]), MIN_COV);  // clamped value + MIN_COV
            ec += max(PK::hi(hist[mec - t]), MIN_COV);
        }
        sc = warp_sum(sc);
        ec = warp_sum(ec);
        if (div == 0) {
            sc = 0;
            ec = 0;
        } else {
            sc /= div;
            ec /= div;
        }
        if (sc >= 10 * ec || ec >= 10 * sc) flags |= kFlagCov;
    } else {
        flags = 0;  // .self.flag is only written with del_telomere (filter.cpp:757-765)
    }

    // ---- final mask (filter.cpp:777-788)
    const int2 q = rd.qvmask[read];
    int2 mk;
    if (P.use_qv_mask && P.use_coverage_mask)
        mk = make_int2(max(maxstart, q.x), min(maxend, q.y));
    else if (P.use_coverage_mask && !P.use_qv_mask)
        mk = make_int2(maxstart, maxend);
    else
        mk = q;

    // ---- optional dump of the profile for .coverage.txt (filter.cpp:599-602)
    if (out.cov0) {
        int* dst = out.cov0 + out.cov0_off[read];
        for (int j = lane; j < L0; j += 32) dst[j] = PK::lo(hist[j]);
    }

    // ---- hinge pre-test: mean coverage near both mask ends (filter.cpp:842-865)
    const int NHR = P.no_hinge_region;
    int cs = 0, ns = 0, ce = 0, ne = 0;
    {
        // bins with  mk.x <= 40 j <= mk.x + NHR
        int jlo = mk.x <= 0 ? 0 : (mk.x + reso - 1) / reso;
        int jhi = mk.x + NHR < 0 ? -1 : min((mk.x + NHR) / reso, L0 - 1);
        for (int j = jlo + lane; j <= jhi; j += 32) {
            cs += PK::lo(hist[j]);
            ns++;
        }
        // bins with  mk.y - NHR <= 40 j <= mk.y
        jlo = mk.y - NHR <= 0 ? 0 : (mk.y - NHR + reso - 1) / reso;
        jhi = mk.y < 0 ? -1 : min(mk.y / reso, L0 - 1);
        for (int j = jlo + lane; j <= jhi; j += 32) {
            ce += PK::lo(hist[j]);
            ne++;
        }
        cs = warp_sum(cs);
        ns = warp_sum(ns);
        ce = warp_sum(ce);
        ne = warp_sum(ne);
    }
    // float on purpose: 0/0 = NaN makes the '< 10' test false (filter.cpp:861-865)
    const float avg_end = __fdiv_rn((float)ce, (float)ne);
    const float avg_start = __fdiv_rn((float)cs, (float)ns);
    const bool skip_hinges = fabsf(__fsub_rn(avg_end, avg_start)) < 10.0f;

    // ---- repeat annotation from the coverage gradient (filter.cpp:796-813);
    // entries are compacted in place over the histogram words they came from
    const int jn = max(L0 - 2, 0);
    int cnt = 0;
    for (int base = 0; base < jn; base += 32) {
        const int j = base + lane;
        int type = 0;
        if (j < jn) {
            const int pos = reso * j;
            if (pos >= mk.x + NHR && pos <= mk.y - NHR) {
                const int c0 = PK::lo(hist[j]);
                const int g = PK::lo(hist[j + 1]) - c0;
                const int thr = min(max((c0 + MIN_COV) / P.coverage_fraction,
                                        P.min_repeat_annotation_threshold),
                                    P.max_repeat_annotation_threshold);
                type = g > thr ? 1 : (g < -thr ? -1 : 0);
            }
        }
        const unsigned am = __ballot_sync(0xffffffffu, type != 0);
        __syncwarp();
        if (type != 0) {
            const int slot = cnt + __popc(am & ((1u << lane) - 1u));
            hist[slot] = (W)(((unsigned)(reso * j) << 2) | (unsigned)(type + 1));
        }
        cnt += __popc(am);
        __syncwarp();
    }

    // ---- merge pass (filter.cpp:817-829), lane 0: +1,+1 closer than the gap
    // threshold drops the later one; -1,-1 drops the earlier one
    int kept = 0;
    if (lane == 0 && cnt > 0) {
        const int GAP = P.repeat_annotation_gap_threshold;
        unsigned cur = (unsigned)hist[0];
        for (int k = 1; k < cnt; k++) {
            const unsigned nxt = (unsigned)hist[k];
            const int ct = (int)(cur & 3u) - 1, nt = (int)(nxt & 3u) - 1;
            const int gap = (int)(nxt >> 2) - (int)(cur >> 2);
            if (ct == 1 && nt == 1 && gap < GAP) {
                continue;  // erase next
            } else if (ct == -1 && nt == -1 && gap < GAP) {
                cur = nxt;  // erase current
            } else {
                hist[kept++] = (W)cur;
                cur = nxt;
            }
        }
        hist[kept++] = (W)cur;
    }
    kept = __shfl_sync(0xffffffffu, kept, 0);
    int off = 0;
    if (lane == 0 && kept > 0) {
        off = atomicAdd(&out.counters[0], kept);
        if (off + kept > out.anno_cap) {
            atomicExch(&out.counters[2], 1);
            off = -1;
        }
    }
    off = __shfl_sync(0xffffffffu, off, 0);
    __syncwarp();
    if (kept > 0 && off >= 0)
        for (int k = lane; k < kept; k += 32) {
            const unsigned w = (unsigned)hist[k];
            out.anno_pool[off + k] = make_int2((int)(w >> 2), (int)(w & 3u) - 1);
            out.hinge_keep[off + k] = 0;
        }
    __syncwarp();  // lane 0 reads the first two annotations back for the work item
    if (lane == 0) {
        store_mask(out, read, mk);
        out.cmask[read] = make_int2(msc, mec);
        out.rflags[read] = flags | (skip_hinges ? kFlagSkipHinge : 0);
        out.anno_ref[read] = make_int2(off, kept);
        if (kept > 0 && !skip_hinges && off >= 0)
            push_work_item(out, read, o0, (int)min((int64_t)0x7fffffff, o1 - o0), mk, off, kept);
    }
    __syncwarp();
}

// Generic path for reads longer than kFlatBins bins or deeper than 32000 records:
// 64-bit packed words in a global scratch slot per warp.
__global__ void __launch_bounds__(128)
k_mask_anno_big(RecView rv, ReadView rd, hg_filter_params P, const int* __restrict__ scal,
                MaskAnnoOut out, unsigned long long* scratch, int slot_words) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int MIN_COV = scal[1];
    const int nbig = out.counters[3];
    for (int w = warp; w < nbig; w += nwarps) {
        const int read = out.big_list[w];
        const int nbz = bins_needed(rd.rlen[read], P);
        mask_anno_read<unsigned long long>(rv, rd, P, MIN_COV, read,
                                           scratch + (size_t)warp * slot_words, nbz, out);
    }
}

// ------------------------------------------------------------------ K4

struct __align__(8) KeyIdx {
    int key, idx;
};
struct GreaterKey {
    __device__ bool operator()(const KeyIdx& a, const KeyIdx& b) const { return a.key > b.key; }
};
struct FirstAsc {
    __device__ bool operator()(const int2& a, const int2& b) const { return a.x < b.x; }
};
struct FirstDesc {
    __device__ bool operator()(const int2& a, const int2& b) const { return a.x > b.x; }
};

// One pile-up record as the hinge call sees it (filter.cpp:877-890): A interval,
// overhangs of B beyond the match w.r.t. B's mask (swapped for complemented
// matches) and the sort key of compare_overlap (LAInterface.cpp:4884).
struct PileRec {
    int as, ae, lo, ro, key;
    bool active;
};
__device__ __forceinline__ PileRec load_pile_rec(const RecView& rv, const ReadView& rd,
                                                 const MaskView& mask, int read, int64_t k) {
    PileRec r;
    const int b = __ldg(rv.bread + k);
    r.active = b != read;  // A == B records are inactive (filter.cpp:538-547)
    r.as = __ldg(rv.abpos + k);
    r.ae = __ldg(rv.aepos + k);
    const int comp = __ldg(rv.flags + k) & 1;
    int bs = __ldg(rv.bbpos + k), be = __ldg(rv.bepos + k);
    if (comp) {  // B to its forward strand, LAInterface.cpp:1619-1626
        const int bl = __ldg(rd.rlen + b);
        const int t = bl - be;
        be = bl - bs;
        bs = t;
    }
    const int2 mb = mask.get(b);
    const int r0 = max(mb.y - be, 0), l0 = max(bs - mb.x, 0);
    r.ro = comp ? l0 : r0;
    r.lo = comp ? r0 : l0;
    r.key = (r.ae - r.as) + (be - bs);
    return r;
}

// Selection test of one record for one annotation (filter.cpp:894-907, 991-1003):
// out-hinge (-1): B keeps going right of its match and ends on A at the
// annotation -> (astart, left overhang); in-hinge (+1) mirrored.
__device__ __forceinline__ bool hinge_select(const PileRec& r, bool out_hinge, int apos, int THETA,
                                             int HTL, int2* e) {
    if (out_hinge) {
        *e = make_int2(r.as, r.lo);
        return r.ro > THETA && r.ae > apos - HTL && r.ae < apos + HTL;
    }
    *e = make_int2(r.ae, r.ro);
    return r.lo > THETA && r.as > apos - HTL && r.as < apos + HTL;
}

// Bridged / unbridged walk over the sorted end list (filter.cpp:916-965,
// 1013-1065); returns true when the annotation is a hinge.
__device__ bool hinge_walk(const int2* ends, int support, bool out_hinge, int2 mk,
                           const hg_filter_params& P) {
    const int THETA = P.theta, HBL = P.hinge_bin_length, U = P.hinge_read_unbridged_threshold;
    bool bridged = true;
    int considered = 0, to_end = 0;
    const int first0 = ends[0].x;
    for (int id = 0; id < support; ++id) {
        const int f = ends[id].x, oh = ends[id].y;
        const int dist_end = out_hinge ? f - mk.x : mk.y - f;
        const int dist0 = out_hinge ? f - first0 : first0 - f;
        if (dist_end < HBL) {
            considered++;
            to_end++;
            if (to_end > U || (considered > U && dist0 > HBL)) {
                bridged = false;
                break;
            }
        } else if (oh < THETA) {
            considered++;
            if (to_end > U || (considered > U && dist0 > HBL)) {
                bridged = false;
                break;
            }
        } else if (oh > THETA) {
            considered++;
            int pl = 1;
            for (int id1 = id + 1; id1 < support; id1++) {
                const int g = out_hinge ? ends[id1].x - f : f - ends[id1].x;
                if (g < HBL)
                    pl++;
                else
                    break;
            }
            if (pl > P.hinge_bin_pileup_threshold) {
                bridged = true;
                break;
            }
        }
    }
    return !bridged && support > P.hinge_min_support;
}

// What the walk can tell apart: reads reaching the mask end (1), small overhang
// (2), internal (3), overhang == theta (0, ignored by the reference's if-chain).
__device__ __forceinline__ int hinge_class(int2 e, bool out_hinge, int2 mk, int THETA, int HBL) {
    const int dist_end = out_hinge ? e.x - mk.x : mk.y - e.x;
    if (dist_end < HBL) return 1;
    return e.y < THETA ? 2 : (e.y > THETA ? 3 : 0);
}

// The same walk, all 32 lanes (call with identical arguments).  Every quantity the
// sequential loop carries is a prefix count over the sorted list, and a pile-up of more
// than T reads inside one bin is just "entry id + T exists and lies within HBL", so each
// lane evaluates the break conditions of its entries and the first entry that fires wins.
__device__ bool hinge_walk_warp(const int2* ends, int support, bool out_hinge, int2 mk,
                                const hg_filter_params& P) {
    const int lane = lane_id();
    const unsigned lt = (1u << lane) - 1u;
    const int THETA = P.theta, HBL = P.hinge_bin_length, U = P.hinge_read_unbridged_threshold;
    const int T = P.hinge_bin_pileup_threshold;
    const int first0 = ends[0].x;
    int considered = 0, to_end = 0;  // running totals before the current chunk
    for (int base = 0; base < support; base += 32) {
        const int id = base + lane;
        int cls = 0, f = 0;
        if (id < support) {
            const int2 e = ends[id];
            f = e.x;
            cls = hinge_class(e, out_hinge, mk, THETA, HBL);
        }
        const unsigned m_cons = __ballot_sync(0xffffffffu, cls != 0);
        const unsigned m_end = __ballot_sync(0xffffffffu, cls == 1);
        const int cons_incl = considered + __popc(m_cons & lt) + (cls != 0 ? 1 : 0);
        const int end_incl = to_end + __popc(m_end & lt) + (cls == 1 ? 1 : 0);
        const int dist0 = out_hinge ? f - first0 : first0 - f;
        bool fire_u = false, fire_b = false;
        if (cls == 1 || cls == 2) {
            fire_u = end_incl > U || (cons_incl > U && dist0 > HBL);
        } else if (cls == 3) {
            // pl = 1 + following entries closer than HBL; pl > T  <=>  entry id + T is that close
            if (T <= 0) {
                fire_b = true;
            } else if (id + T < support) {
                const int f2 = ends[id + T].x;
                fire_b = (out_hinge ? f2 - f : f - f2) < HBL;
            }
        }
        const unsigned m_fire = __ballot_sync(0xffffffffu, fire_u || fire_b);
        if (m_fire) {
            const int src = __ffs(m_fire) - 1;
            const bool bridged = __shfl_sync(0xffffffffu, fire_b ? 1 : 0, src) != 0;
            return !bridged && support > P.hinge_min_support;
        }
        considered += __popc(m_cons);
        to_end += __popc(m_end);
    }
    return false;  // never left the initial `bridged = true`
}

// One warp per annotated read.  The reference feeds the walk with the selected
// records in the order left by two unstable std::sorts (pile-up by length, then
// the end list by position).  The walk only distinguishes entries by (position,
// class), so whenever no two selected entries share a position with different
// classes ANY sort by position yields the reference's result: that is the fast
// path (warp rank sort, no pile-up sort at all).  Otherwise the annotation goes
// through the order-exact path: libstdc++'s introsort restated (hg_order.h) on
// the pile-up and on the end list.
//
// Scratch slot per warp (cap = deepest pile-up):
//   int4 rec[cap] | KeyIdx ord[cap] | int2 ends[cap] | int2 sorted[cap] | int2 keys[cap] | int gl[2 cap]
constexpr int kHingeSlotBytesPerRec = 16 + 8 + 8 + 8 + 8 + 8;
constexpr int kHingeSmemEnds = 192;

__global__ void __launch_bounds__(128, 8)
k_hinge_call(RecView rv, ReadView rd, hg_filter_params P, MaskView mask,
             const int2* __restrict__ anno_ref, const int2* __restrict__ anno_pool,
             int* __restrict__ counters, const int4* __restrict__ work_items, int* __restrict__ exact_list,
             uint8_t* __restrict__ hinge_keep, uint8_t* scratch, int cap, int4* __restrict__ item_log,
             PeerView pv) {
    if (pv.world > 1) {
        // sharded run: the masks of the other ranks' reads arrive through peer memory (store_mask);
        // every CTA waits on the local arrival flags before its first look-up
        if (threadIdx.x < pv.world) {
            if (!peer_wait(pv.flag_mask(pv.rank) + threadIdx.x, pv.epoch)) atomicExch(&counters[15], 1);
            // a rank whose annotation pool overflowed makes every rank rerun (hg_filter's retry)
            if (ld_acquire_sys(pv.flag_ovf(pv.rank) + threadIdx.x) == pv.epoch) atomicExch(&counters[2], 1);
        }
        __syncthreads();
    }
    const int lane = lane_id();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    uint8_t* base = scratch + (size_t)warp * cap * kHingeSlotBytesPerRec;
    int2* ends = reinterpret_cast<int2*>(base + (size_t)cap * 24);
    int2* sorted = reinterpret_cast<int2*>(base + (size_t)cap * 32);
    int2* keys = reinterpret_cast<int2*>(base + (size_t)cap * 40);  // (total length, selection index)
    int* const gl = reinterpret_cast<int*>(base + (size_t)cap * 48);  // warp_sort_exact position lists
    const int nwork = counters[1];
    const int THETA = P.theta, HTL = P.hinge_tolerance_length, HBL = P.hinge_bin_length;
    // end lists of up to kHingeSmemEnds entries are sorted and walked in shared memory
    __shared__ int2 sh_ends[4][kHingeSmemEnds], sh_sorted[4][kHingeSmemEnds];
    __shared__ __align__(16) int sh_keys[4][kHingeSmemEnds];
    int2* const ends_s = sh_ends[threadIdx.x >> 5];
    int2* const sorted_s = sh_sorted[threadIdx.x >> 5];
    int* const keys_s = sh_keys[threadIdx.x >> 5];
    (void)nwarps;

    // reads are handed out one at a time: a read that needs the (sequential) order-exact
    // path keeps its warp busy for a long time and must not delay a static share of others
    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(&counters[5], 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= nwork) break;
        const long long t_begin = item_log ? clock64() : 0;
        int log_support = 0, log_exact = 0;
        const int4 it0 = __ldg(work_items + 3 * (size_t)w), it1 = __ldg(work_items + 3 * (size_t)w + 1);
        const int4 it2 = __ldg(work_items + 3 * (size_t)w + 2);
        const int read = it0.x;
        const int64_t o0 = (int64_t)(((unsigned long long)(unsigned)it0.w << 32) | (unsigned)it0.z);
        const int64_t o1 = o0 + it0.y;
        const int2 mk = make_int2(it1.x, it1.y);
        const int2 ar = make_int2(it1.z, it1.w);
        bool deferred = false;
        for (int j = 0; j < ar.y; j++) {
            const int2 an = j == 0 ? make_int2(it2.x, it2.y) : (j == 1 ? make_int2(it2.z, it2.w) : anno_pool[ar.x + j]);
            const int apos = an.x;
            const bool out_hinge = an.y == -1;
            // ---- select in file order, two steps: (1) the test on A's coordinates alone
            // (two coalesced columns) keeps a minority of the pile-up; (2) only those pay for
            // the B side (four more columns + gathers of rlen[B] and mask[B])
            int* const near_idx = gl;  // free here: the position lists are only live inside a sort
            int n_near = 0;
            const int np = (int)(o1 - o0);
            const int* __restrict__ pcol = (out_hinge ? rv.aepos : rv.abpos) + o0;
            for (int kb = 0; kb < np; kb += 128) {  // four independent loads in flight per lane
                int pos[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int k = kb + 32 * u + lane;
                    pos[u] = k < np ? __ldg(pcol + k) : apos + HTL;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const bool near = pos[u] > apos - HTL && pos[u] < apos + HTL;
                    const unsigned nm = __ballot_sync(0xffffffffu, near);
                    if (near) near_idx[n_near + __popc(nm & ((1u << lane) - 1u))] = kb + 32 * u + lane;
                    n_near += __popc(nm);
                }
            }
            __syncwarp();
            int support = 0;
            for (int cb = 0; cb < n_near; cb += 32) {
                const int c = cb + lane;
                bool sel = false;
                int2 e = make_int2(0, 0);
                int sel_key = 0;
                if (c < n_near) {
                    const PileRec r = load_pile_rec(rv, rd, mask, read, o0 + near_idx[c]);
                    sel = r.active && hinge_select(r, out_hinge, apos, THETA, HTL, &e);
                    sel_key = r.key;
                }
                const unsigned sm = __ballot_sync(0xffffffffu, sel);
                if (sel) {
                    const int slot = support + __popc(sm & ((1u << lane) - 1u));
                    ends[slot] = e;
                    keys[slot] = make_int2(sel_key, slot);
                    if (slot < kHingeSmemEnds) ends_s[slot] = e;
                }
                support += __popc(sm);
            }
            __syncwarp();
            uint8_t keep = 0;
            if (support >= P.hinge_min_support) {  // filter.cpp:910, 1005
                const bool small = support <= kHingeSmemEnds;
                const int2* const src = small ? ends_s : ends;
                int2* const dst = small ? sorted_s : sorted;
                // ---- stable rank sort by position (ascending for out-hinges, descending for
                // in-hinges) and detection of position ties the walk could tell apart
                bool danger = support > 1024;
                if (!danger) {
                    // descending order = ascending order of the negated position
                    const int sgn = out_hinge ? 1 : -1;
                    if (small && rd.rlen[read] < (1 << 22)) {
                        // one distinct integer per entry, ordered like (position, index): the rank is a
                        // count of smaller keys, four per 128-bit shared-memory load
                        const int padded = (support + 3) & ~3;
                        for (int i = lane; i < padded; i += 32)
                            keys_s[i] = i < support ? sgn * ends_s[i].x * 256 + i : 0x7fffffff;
                        __syncwarp();
                        for (int i = lane; i < support; i += 32) {
                            const int me = keys_s[i];
                            int rank = 0;
                            for (int t = 0; t < padded; t += 4) {
                                const int4 k4 = *reinterpret_cast<const int4*>(keys_s + t);
                                rank += (k4.x < me) + (k4.y < me) + (k4.z < me) + (k4.w < me);
                            }
                            sorted_s[rank] = ends_s[i];
                        }
                    } else
                    for (int i = lane; i < support; i += 32) {
                        const int2 me = src[i];
                        const int mx = sgn * me.x;
                        int rank = 0;
#pragma unroll 4
                        for (int t = 0; t < support; t++) {
                            const int x = sgn * src[t].x;
                            rank += (x < mx || (x == mx && t < i)) ? 1 : 0;
                        }
                        dst[rank] = me;
                    }
                    __syncwarp();
                    bool d = false;
                    for (int i = lane; i + 1 < support; i += 32) {
                        const int2 a = dst[i], b = dst[i + 1];
                        d = d || (a.x == b.x && hinge_class(a, out_hinge, mk, THETA, HBL) !=
                                                    hinge_class(b, out_hinge, mk, THETA, HBL));
                    }
                    danger = __any_sync(0xffffffffu, d);
                }
                if (!danger) {
                    keep = hinge_walk_warp(dst, support, out_hinge, mk, P) ? 1 : 0;
                } else if (![&]() {
                               // ---- order-exact path, cheap form.  The end list enters its std::sort in
                               // pile-up order = total length, descending.  If the selected records have
                               // pairwise different lengths that order is unique whatever the unstable
                               // pile-up sort did to ties elsewhere: no pile-up sort needed at all.
                               if (support > 1024) return false;
                               bool tie = false;
                               for (int i = lane; i < support; i += 32) {
                                   const int2 me = keys[i];
                                   int rank = 0;
                                   for (int t = 0; t < support; t++) {
                                       const int x = keys[t].x;
                                       rank += (x > me.x || (x == me.x && t < i)) ? 1 : 0;
                                       tie = tie || (x == me.x && t != i);
                                   }
                                   sorted[rank] = ends[me.y];
                               }
                               if (__any_sync(0xffffffffu, tie)) return false;
                               __syncwarp();
                               if (out_hinge)  // filter.cpp:914 / 1010, all lanes
                                   warp_sort_exact(sorted, support, FirstAsc(), gl, gl + cap, keys);
                               else
                                   warp_sort_exact(sorted, support, FirstDesc(), gl, gl + cap, keys);
                               keep = hinge_walk_warp(sorted, support, out_hinge, mk, P) ? 1 : 0;
                               if (lane == 0) atomicAdd(&counters[4], 1);
                               return true;
                           }()) {
                    // ---- order-exact path, full form: the whole pile-up has to go through std::sort.
                    // That is a long, latency-bound job for one warp (~150 us when its scratch is in
                    // global memory) and would set the duration of this kernel, so the read is handed
                    // to k_hinge_exact, which redoes all its annotations out of shared memory.
                    deferred = true;
                }
            }
            if (deferred) break;
            if (lane == 0) hinge_keep[ar.x + j] = keep;
            log_support = max(log_support, support);
            __syncwarp();
        }
        if (deferred && lane == 0) exact_list[atomicAdd(&counters[6], 1)] = read;
        log_exact = deferred ? 1 : 0;
        if (item_log && lane == 0)
            item_log[w] = make_int4(read, (int)(clock64() - t_begin), log_support, log_exact);
    }
}

// Reads whose annotations need the reference's exact sort order (see k_hinge_call): one CTA
// per read, the reference's own sequence for every annotation of the read -- pile-up in file
// order, std::sort by total length (filter.cpp:565-567), selection in that order, std::sort of the
// end list by position (filter.cpp:914 / 1010), walk -- with libstdc++'s introsort restated in
// hg_order.h.  The sort is a chain of short dependent steps, so its arrays live in shared memory
// (48 B per pile-up record); deeper pile-ups than `scap` fall back to a global slot.
constexpr int kHingeExactBytesPerRec = 16 + 8 + 8 + 8 + 8;

__global__ void __launch_bounds__(128)
k_hinge_exact(RecView rv, ReadView rd, hg_filter_params P, MaskView mask,
              const int2* __restrict__ anno_ref, const int2* __restrict__ anno_pool,
              int* __restrict__ counters, int queue_slot, const int* __restrict__ exact_list,
              uint8_t* __restrict__ hinge_keep, uint8_t* gscratch, int gcap, int scap, int np_above, int np_upto) {
    extern __shared__ __align__(16) uint8_t sm_exact[];
    __shared__ CtaSortState sort_state;
    __shared__ int sh_w, sh_n;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int nlist = counters[6];
    const int THETA = P.theta, HTL = P.hinge_tolerance_length;
    for (;;) {
        if (threadIdx.x == 0) sh_w = atomicAdd(&counters[queue_slot], 1);
        __syncthreads();
        const int w = sh_w;
        __syncthreads();
        if (w >= nlist) break;
        const int read = exact_list[w];
        const int64_t o0 = rv.read_off[read], o1 = rv.read_off[read + 1];
        const int np = (int)(o1 - o0);
        if (np <= np_above || np > np_upto) continue;  // another size tier's read (uniform over the CTA)
        const bool in_smem = np <= scap;
        const int cap = in_smem ? scap : gcap;
        uint8_t* base = in_smem ? sm_exact : gscratch + (size_t)blockIdx.x * gcap * kHingeSlotBytesPerRec;
        int4* rec = reinterpret_cast<int4*>(base);
        KeyIdx* ord = reinterpret_cast<KeyIdx*>(base + (size_t)cap * 16);
        int2* ends = reinterpret_cast<int2*>(base + (size_t)cap * 24);
        int2* tmp = reinterpret_cast<int2*>(base + (size_t)cap * 32);
        int* gl = reinterpret_cast<int*>(base + (size_t)cap * 40);
        const int2 mk = mask.full[read];  // the read's own mask: always in the local array
        const int2 ar = anno_ref[read];

        // pile-up in file order: all threads fetch (records, then the gathers of rlen[B] and mask[B]),
        // one warp squeezes out the inactive A == B records (filter.cpp:538-547) in place
        for (int k = threadIdx.x; k < np; k += blockDim.x) {
            const PileRec r = load_pile_rec(rv, rd, mask, read, o0 + k);
            rec[k] = make_int4(r.as, r.ae, r.lo, r.ro);
            ord[k].key = r.key;
            ord[k].idx = r.active ? 1 : 0;
        }
        __syncthreads();
        if (warp == 0) {
            int n = 0;
            for (int kb = 0; kb < np; kb += 32) {
                const int k = kb + lane;
                int4 q = make_int4(0, 0, 0, 0);
                int key = 0;
                bool active = false;
                if (k < np) {
                    q = rec[k];
                    key = ord[k].key;
                    active = ord[k].idx != 0;
                }
                const unsigned am = __ballot_sync(0xffffffffu, active);
                if (active) {
                    const int s = n + __popc(am & lt);
                    rec[s] = q;
                    ord[s].key = key;
                    ord[s].idx = s;
                }
                n += __popc(am);
                __syncwarp();
            }
            if (lane == 0) sh_n = n;
        }
        __syncthreads();
        const int n = sh_n;
        // std::sort by total length (filter.cpp:565-567), the warps on different sub-ranges
        cta_sort_exact(ord, n, GreaterKey(), gl, gl + cap, reinterpret_cast<KeyIdx*>(tmp), &sort_state);

        for (int j = 0; j < ar.y; j++) {
            const int2 an = anno_pool[ar.x + j];
            const bool out_hinge = an.y == -1;
            if (warp == 0) {  // selection in pile-up order
                int support = 0;
                for (int kb = 0; kb < n; kb += 32) {
                    const int k = kb + lane;
                    bool sel = false;
                    int2 e = make_int2(0, 0);
                    if (k < n) {
                        const int4 q = rec[ord[k].idx];
                        PileRec r;
                        r.as = q.x; r.ae = q.y; r.lo = q.z; r.ro = q.w; r.key = 0; r.active = true;
                        sel = hinge_select(r, out_hinge, an.x, THETA, HTL, &e);
                    }
                    const unsigned sm = __ballot_sync(0xffffffffu, sel);
                    if (sel) ends[support + __popc(sm & lt)] = e;
                    support += __popc(sm);
                }
                if (lane == 0) sh_n = support;
            }
            __syncthreads();
            const int support = sh_n;
            uint8_t keep = 0;
            if (support >= P.hinge_min_support) {  // filter.cpp:910, 1005 (uniform over the CTA)
                if (out_hinge)  // filter.cpp:914 / 1010
                    cta_sort_exact(ends, support, FirstAsc(), gl, gl + cap, tmp, &sort_state);
                else
                    cta_sort_exact(ends, support, FirstDesc(), gl, gl + cap, tmp, &sort_state);
                if (warp == 0) {
                    keep = hinge_walk_warp(ends, support, out_hinge, mk, P) ? 1 : 0;
                    if (lane == 0) atomicAdd(&counters[4], 1);
                }
            }
            if (threadIdx.x == 0) hinge_keep[ar.x + j] = keep;
            __syncthreads();
        }
    }
}

// The same for pile-ups of at most `cap` records: ONE WARP per read, four reads per CTA.  A pile-up
// of a few hundred records is far too small for a CTA (k_hinge_exact above spends its time in
// barriers and in the lock of its work stack: ~270 us per read, 296 reads at a time -- 3.4 ms for the
// 20 k reads of the long-read / fragmented-alignment set); a warp takes 50-100 us, most of it the
// fixed cost of the ~n/8 partition steps of the introsort emulation (latency, not throughput), so
// what counts is how many warps are in flight.  Shared memory per record is therefore kept to 28
// bytes -- sort keys, end list, sort scratch and 16-bit position lists; the records themselves sit
// in the warp's slot of the global scratch k_hinge_call has finished with (read once per
// annotation, L1-resident) -- which allows 32 / 21 / 10 warps per SM for the three size tiers.
// Handles the reads of the list with np_above < pile-up size <= cap; launched once per size tier.
constexpr int kHingeWarpBytesPerRec = 8 + 8 + 8 + 4;

__global__ void __launch_bounds__(128)
k_hinge_exact_warp(RecView rv, ReadView rd, hg_filter_params P, MaskView mask,
                   const int2* __restrict__ anno_ref, const int2* __restrict__ anno_pool,
                   int* __restrict__ counters, int queue_slot, const int* __restrict__ exact_list,
                   uint8_t* __restrict__ hinge_keep, uint8_t* gscratch, int gcap, int np_above, int cap) {
    extern __shared__ __align__(16) uint8_t sm_exact[];
    const int lane = lane_id();
    const unsigned lt = (1u << lane) - 1u;
    uint8_t* base = sm_exact + (size_t)(threadIdx.x >> 5) * cap * kHingeWarpBytesPerRec;
    KeyIdx* ord = reinterpret_cast<KeyIdx*>(base);
    int2* ends = reinterpret_cast<int2*>(base + (size_t)cap * 8);
    int2* tmp = reinterpret_cast<int2*>(base + (size_t)cap * 16);
    uint16_t* gl = reinterpret_cast<uint16_t*>(base + (size_t)cap * 24);
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int4* rec = reinterpret_cast<int4*>(gscratch + (size_t)gwarp * gcap * sizeof(int4));
    const int nlist = counters[6];
    const int THETA = P.theta, HTL = P.hinge_tolerance_length;
    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(&counters[queue_slot], 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= nlist) break;
        const int read = exact_list[w];
        const int64_t o0 = rv.read_off[read], o1 = rv.read_off[read + 1];
        const int np = (int)(o1 - o0);
        if (np <= np_above || np > cap) continue;
        const int2 mk = mask.full[read];  // the read's own mask: always in the local array
        const int2 ar = anno_ref[read];
        // pile-up in file order without the inactive A == B records (filter.cpp:538-547)
        int n = 0;
        for (int kb = 0; kb < np; kb += 32) {
            const int k = kb + lane;
            PileRec r;
            r.active = false;
            if (k < np) r = load_pile_rec(rv, rd, mask, read, o0 + k);
            const unsigned am = __ballot_sync(0xffffffffu, r.active);
            if (r.active) {
                const int s = n + __popc(am & lt);
                rec[s] = make_int4(r.as, r.ae, r.lo, r.ro);
                ord[s].key = r.key;
                ord[s].idx = s;
            }
            n += __popc(am);
        }
        __syncwarp();
        // std::sort by total length (filter.cpp:565-567)
        warp_sort_exact(ord, n, GreaterKey(), gl, gl + cap, reinterpret_cast<KeyIdx*>(tmp));
        for (int j = 0; j < ar.y; j++) {
            const int2 an = anno_pool[ar.x + j];
            const bool out_hinge = an.y == -1;
            int support = 0;
            for (int kb = 0; kb < n; kb += 32) {  // selection in pile-up order
                const int k = kb + lane;
                bool sel = false;
                int2 e = make_int2(0, 0);
                if (k < n) {
                    const int4 q = rec[ord[k].idx];
                    PileRec r;
                    r.as = q.x; r.ae = q.y; r.lo = q.z; r.ro = q.w; r.key = 0; r.active = true;
                    sel = hinge_select(r, out_hinge, an.x, THETA, HTL, &e);
                }
                const unsigned sm = __ballot_sync(0xffffffffu, sel);
                if (sel) ends[support + __popc(sm & lt)] = e;
                support += __popc(sm);
            }
            __syncwarp();
            uint8_t keep = 0;
            if (support >= P.hinge_min_support) {  // filter.cpp:910, 1005
                if (out_hinge)  // filter.cpp:914 / 1010
                    warp_sort_exact(ends, support, FirstAsc(), gl, gl + cap, tmp);
                else
                    warp_sort_exact(ends, support, FirstDesc(), gl, gl + cap, tmp);
                keep = hinge_walk_warp(ends, support, out_hinge, mk, P) ? 1 : 0;
                if (lane == 0) atomicAdd(&counters[4], 1);
            }
            if (lane == 0) hinge_keep[ar.x + j] = keep;
            __syncwarp();
        }
    }
}

// Test hook: sorts `count` independent arrays of (key, idx) pairs with warp_sort_exact.
__global__ void k_debug_warp_sort(KeyIdx* data, const int* off, int count, int descending, int* g, int* l,
                                  KeyIdx* tmp) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= count) return;
    const int o = off[w], n = off[w + 1] - o;
    if (descending)
        warp_sort_exact(data + o, n, GreaterKey(), g + o, l + o, tmp + o);
    else
        warp_sort_exact(data + o, n, [] __device__(const KeyIdx& a, const KeyIdx& b) { return a.key < b.key; },
                        g + o, l + o, tmp + o);
}

// The same test for cta_sort_exact: one CTA of 128 threads per array.
__global__ void __launch_bounds__(128)
k_debug_cta_sort(KeyIdx* data, const int* off, int count, int descending, int* g, int* l, KeyIdx* tmp) {
    __shared__ CtaSortState st;
    const int w = blockIdx.x;
    if (w >= count) return;
    const int o = off[w], n = off[w + 1] - o;
    if (descending)
        cta_sort_exact(data + o, n, GreaterKey(), g + o, l + o, tmp + o, &st);
    else
        cta_sort_exact(data + o, n, [] __device__(const KeyIdx& a, const KeyIdx& b) { return a.key < b.key; },
                       g + o, l + o, tmp + o, &st);
}

void launch_debug_warp_sort(void* data, const int* off, int count, int descending, int* g, int* l, void* tmp,
                            cudaStream_t st) {
    if (getenv("HINGE_B200_DEBUG_SORT_CTA")) {
        k_debug_cta_sort<<<count, 128, 0, st>>>((KeyIdx*)data, off, count, descending, g, l, (KeyIdx*)tmp);
        return;
    }
    k_debug_warp_sort<<<(count * 32 + 127) / 128, 128, 0, st>>>((KeyIdx*)data, off, count, descending, g, l,
                                                                (KeyIdx*)tmp);
}

__global__ void k_max_pileup(const int64_t* __restrict__ read_off, int n_read, int* out_max) {
    int m = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_read; i += gridDim.x * blockDim.x)
        m = max(m, (int)min((int64_t)0x7fffffff, read_off[i + 1] - read_off[i]));
    m = warp_max(m);
    if (lane_id() == 0 && m > 0) atomicMax(out_max, m);
}

// ------------------------------------------------------------------ launchers

static inline int ceil_div64(int64_t a, int b) { return (int)((a + b - 1) / b); }

void launch_csr_validate(const RecView& rv, const ReadView& rd, int64_t* read_off, int* self_cnt, int* err,
                         cudaStream_t st) {
    cudaMemsetAsync(self_cnt, 0, sizeof(int) * rd.n_read, st);
    k_csr_validate<<<ceil_div64(rv.novl + 1, 256), 256, 0, st>>>(rv, rd, read_off, self_cnt, err);
    g_launches += 1;
}

void launch_qv_mask(int n_read, const int64_t* qv_off, const uint8_t* qv, int tspace, int2* out,
                    cudaStream_t st) {
    k_qv_mask<<<ceil_div64(n_read, 128), 128, 0, st>>>(n_read, qv_off, qv, tspace, out);
    g_launches += 1;
}

// mode 0: accumulate + pick + radix fallback in one go (single context)
// mode 1: accumulate the owned reads only (sharded, before the all-reduce of the histogram)
// mode 2: pick from the summed histogram (sharded, after it)
void launch_median(const ReadView& rd, const hg_filter_params& P, FilterScratch& s, int mode,
                   cudaStream_t st) {
    // med_hist is zero on entry (cleared at allocation / binding and by every run's pick)
    if (mode == 2) {
        g_launches += 1;
        k_median_pick<<<1, 256, 0, st>>>(s.med_hist, P.est_cov, P.min_cov, s.scal);
        return;
    }
    g_launches += mode == 0 ? 2 : 1;
    k_median_hist<<<148 * 2, 256, 0, st>>>(s.mean_cov, mode == 0 ? 0 : rd.r_lo, mode == 0 ? rd.n_read : rd.r_hi,
                                           s.med_hist, P.est_cov, P.min_cov, s.scal, mode == 0);
    if (mode == 0) k_median_radix<<<1, 1024, 0, st>>>(s.mean_cov, rd.n_read, P.est_cov, P.min_cov, s.scal);
}

void launch_peer_hist_push(FilterScratch& s, const PeerView& peer, cudaStream_t st) {
    g_launches += 1;
    k_peer_hist_push<<<1, 1024, 0, st>>>(s.med_hist, peer);
}

void launch_peer_median_pick(const hg_filter_params& P, FilterScratch& s, const PeerView& peer, cudaStream_t st) {
    g_launches += 1;
    k_peer_median_pick<<<1, 256, 0, st>>>(peer, P.est_cov, P.min_cov, s.scal, s.counters + 15);
}

void launch_peer_signal_masks(FilterScratch& s, const PeerView& peer, cudaStream_t st) {
    g_launches += 1;
    k_peer_signal_masks<<<1, 32, 0, st>>>(peer, s.counters);
}

void launch_mask_anno(const RecView& rv, const ReadView& rd, const hg_filter_params& P,
                      int r_begin, int r_end, FilterScratch& s, int* cov0, const int64_t* cov0_off,
                      const PeerView& peer, cudaStream_t st) {
    MaskAnnoOut out;
    out.peer = peer;
    out.mask = s.mask; out.mask_pk = s.mask_pk; out.mask_g = s.mask_g; out.cmask = s.cmask; out.rflags = s.rflags; out.anno_ref = s.anno_ref;
    out.anno_pool = s.anno_pool; out.hinge_keep = s.hinge_keep; out.anno_cap = s.anno_cap; out.counters = s.counters;
    out.work_items = s.work_items; out.big_list = s.big_list; out.cov0 = cov0; out.cov0_off = cov0_off;
    // no clearing here: every read of the planned range gets its results written (flat or generic
    // path), the rest was cleared when the plan was made; the counters are zeroed by launch_profile
    launch_mask_anno_flat(rv, rd, P, r_begin, r_end, s, out, st);
    g_launches += (s.big_slot_words > 0);
    if (s.big_slot_words > 0)
        k_mask_anno_big<<<s.big_warps / 4, 128, 0, st>>>(rv, rd, P, s.scal, out, s.big_scratch,
                                                         s.big_slot_words);
}

void launch_max_pileup(const int64_t* read_off, int n_read, int* out_max, cudaStream_t st) {
    cudaMemsetAsync(out_max, 0, sizeof(int), st);
    k_max_pileup<<<148 * 2, 256, 0, st>>>(read_off, n_read, out_max);
    g_launches += 1;
}

void launch_hinge_call(const RecView& rv, const ReadView& rd, const hg_filter_params& P,
                       FilterScratch& s, const PeerView& peer, cudaStream_t st) {
    if (s.hinge_cap <= 0) return;
    g_launches += 2;
    MaskView mv;
    mv.full = s.mask;
    mv.packed = s.mask_pk;
    mv.g = s.mask_g;
    k_hinge_call<<<s.hinge_warps / 4, 128, 0, st>>>(rv, rd, P, mv, s.anno_ref, s.anno_pool,
                                                    s.counters, s.work_items, s.exact_list, s.hinge_keep,
                                                    s.hinge_scratch, s.hinge_cap, s.item_log, peer);
    // The reads that need the exact sort order, in size tiers that run side by side on forked streams
    // (they work on different reads of one list; each has its own queue cursor).  One WARP per read
    // (k_hinge_exact_warp) takes everything that fits a warp's slice of shared memory:
    //   A   <= 256 records: 7 KB per warp, 32 warps per SM        B   <= 384: 10.5 KB, 21 warps per SM
    //   C   <= 768: 21 KB, 10 warps per SM
    //   D   deeper: one CTA per read (sub-ranges of the introsort on different warps), 72 KB, global
    //       scratch beyond 1536 records
    constexpr int capA = 256, capB = 384, capC = 768, capD = 1536;
    const int smemD = capD * kHingeExactBytesPerRec;
    // function attributes are per device: set them on every launch (cheap), not once per process
    cudaFuncSetAttribute(k_hinge_exact_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * capC * kHingeWarpBytesPerRec);
    cudaFuncSetAttribute(k_hinge_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, smemD);
    g_launches += 4;
    const bool fork = s.side_stream[0] != nullptr;
    cudaStream_t sB = fork ? s.side_stream[0] : st, sC = fork ? s.side_stream[1] : st, sD = fork ? s.side_stream[2] : st;
    if (fork) {
        cudaEventRecord(s.side_event[0], st);
        cudaStreamWaitEvent(sB, s.side_event[0], 0);
        cudaStreamWaitEvent(sC, s.side_event[0], 0);
        cudaStreamWaitEvent(sD, s.side_event[0], 0);
    }
    // The global scratch k_hinge_call has finished with (hinge_warps slots of 56 B per record) is
    // carved up again: tier D's CTAs keep full slots, a warp of the other tiers only needs 16 B per
    // record for the records themselves.
    const size_t total = (size_t)s.hinge_warps * s.hinge_cap * kHingeSlotBytesPerRec;
    const size_t slotD = (size_t)s.hinge_cap * kHingeSlotBytesPerRec, slotW = (size_t)s.hinge_cap * sizeof(int4) * 4;
    const int gridD = (int)std::max<size_t>(1, std::min<size_t>(2 * s.num_sms, total / 4 / slotD));
    const size_t ctas = (total - slotD * gridD) / slotW;  // CTAs of four warps the rest can serve
    const int gA = (int)std::max<size_t>(1, std::min<size_t>(8 * s.num_sms, ctas / 2));
    const int gB = (int)std::max<size_t>(1, std::min<size_t>(5 * s.num_sms, ctas / 4));
    const int gC = (int)std::max<size_t>(1, std::min<size_t>(2 * s.num_sms, ctas / 8));
    uint8_t* const baseW = s.hinge_scratch + slotD * gridD;
    k_hinge_exact_warp<<<gA, 128, 4 * capA * kHingeWarpBytesPerRec, st>>>(
        rv, rd, P, mv, s.anno_ref, s.anno_pool, s.counters, 12, s.exact_list, s.hinge_keep, baseW, s.hinge_cap, 0, capA);
    k_hinge_exact_warp<<<gB, 128, 4 * capB * kHingeWarpBytesPerRec, sB>>>(
        rv, rd, P, mv, s.anno_ref, s.anno_pool, s.counters, 13, s.exact_list, s.hinge_keep, baseW + slotW * gA,
        s.hinge_cap, capA, capB);
    k_hinge_exact_warp<<<gC, 128, 4 * capC * kHingeWarpBytesPerRec, sC>>>(
        rv, rd, P, mv, s.anno_ref, s.anno_pool, s.counters, 14, s.exact_list, s.hinge_keep, baseW + slotW * (gA + gB),
        s.hinge_cap, capB, capC);
    k_hinge_exact<<<gridD, 128, smemD, sD>>>(rv, rd, P, mv, s.anno_ref, s.anno_pool, s.counters, 7, s.exact_list,
                                            s.hinge_keep, s.hinge_scratch, s.hinge_cap, capD, capC, 0x7fffffff);
    if (fork) {
        for (int i = 0; i < 3; i++) {
            cudaEventRecord(s.side_event[1 + i], i == 0 ? sB : (i == 1 ? sC : sD));
            cudaStreamWaitEvent(st, s.side_event[1 + i], 0);
        }
    }
}

}  // namespace hg
