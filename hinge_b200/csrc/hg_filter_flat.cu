// The pile-up scan of `hinge filter` (/root/reference/src/filter/filter.cpp:588-865,
// /root/reference/src/lib/LAInterface.cpp:4298-4320) in flat form: a BATCH of consecutive
// A-reads per CTA instead of one read per warp, and the records read from HBM ONCE.
//
// Why this shape.  A PacBio-like read has ~90 coverage bins and ~90 pile-up records: far too
// little work for a warp, so a warp-per-read kernel spends its time in per-read fixed
// overhead executed by 32 mostly idle lanes (first cut of this round, ncu: 606 warp
// instructions per read, 65 % issue-active, 11 % of HBM).  Here the profiles of ~40 reads are
// laid end to end in one shared-memory array and every phase runs flat over it with all lanes
// busy.  The stage has one global dependency -- MIN_COV, from the median of the per-read mean
// coverage (filter.cpp:642-678) -- so it is split there:
//
//   K1 k_profile_flat    (before the median) builds both coverage profiles of every read:
//     scatter   flat over the batch's records (int4 loads of aread/abpos/aepos, the next
//               step's loads in flight while this step's events go out): four packed +-1
//               events per record (low half: profile without cut-off, high half: with).
//               The lanes of a warp are spread over eight record windows (flat_group) so
//               that the many records that start in their read's first bin or end in its
//               last one do not all serialise on one shared-memory word.
//     scan      ONE block-wide prefix sum over the concatenated array.  Every record adds
//               +1 and -1 inside its own read's bins, so the running sum is back at zero at
//               every read boundary: no segmentation needed.  The scanned words go to HBM
//               (4 B per bin, about 4 B per record) for K2.
//     per read  profile length and mean coverage (filter.cpp:642-656), self-overlap flag.
//   K2 k_mask_anno_flat  (after the median) streams the stored profiles: bit maps of bins whose
//     cut-off coverage is <= MIN_COV ("zeros") and of bins where the coverage jumps by more
//     than the smallest annotation threshold, then one thread per read walks its slice of the
//     maps: longest covered run (filter.cpp:696-728), mask, telomere flag, repeat annotations
//     with the streaming form of the merge pass (filter.cpp:796-829), hinge pre-test
//     (filter.cpp:842-865).
//
// K1 is bound by the shared-memory data pipe (ncu: l1tex data-pipe wavefronts > 70 % of peak,
// ~4 wavefronts per ATOMS on random bins is what the banks give), so everything else in it is
// kept conflict-free (sw) and the global accesses wide.  Tried and measured on B200, K2 of the
// previous form alone: match.any aggregation of equal bins 0.84 ms, ballot aggregation of the
// hot bins 0.54 ms, none 0.40 ms, + lane spreading 0.365 ms, + swizzle and prefetch 0.34 ms.
//
// Reads longer than kFlatBins bins and pile-ups deeper than the 16-bit halves can count go
// to per-read fallbacks (k_cov_big here, k_mask_anno_big in hg_filter.cu); so does everything
// when MIN_COV < 0 (covered runs could then cross read boundaries).
#include "hg_device.cuh"
#include "hg_filter.h"

namespace hg {

extern int64_t g_launches;

constexpr int kFlatThreads = 256;
constexpr int kFlatItems = 16;                           // bins per thread and scan pass
constexpr int kFlatPass = kFlatThreads * kFlatItems;     // bins per scan pass
constexpr int kFlatWords = kFlatBins + 32;               // histogram words per CTA (+ slack)
constexpr int kFlatMaps = (kFlatBins + kFlatPass - 1) / kFlatPass * kFlatThreads;  // 16-bit map entries

__device__ __forceinline__ uint4 lds128(const uint32_t* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void sts128(uint32_t* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }
__device__ __forceinline__ int f_lo(uint32_t v) { return (int)(v & 0xffffu); }   // after the scan: cov0 >= 0
__device__ __forceinline__ int f_hi(uint32_t v) { return (int)v >> 16; }

// Shared-memory layout of the histogram.  A thread of the scan owns 16 consecutive words and
// moves them as four 128-bit vectors; with the identity layout the 8 lanes of a quarter warp
// start 64 B apart and hit only two of the eight 16-byte bank groups (ncu: twice the ideal
// wavefronts on every LDS.128 / STS.128).  Flipping word-index bits 2-3 with bits 5-6 keeps
// every aligned 4-word group intact and makes those accesses conflict-free.
__device__ __forceinline__ int sw(int j) { return j ^ ((j >> 3) & 12); }

// Bit `i` of the map <=> entry i (maps are arrays of 32-bit words in shared memory).
__device__ __forceinline__ uint32_t map_word(const uint32_t* map, int w, int lo, int hi) {
    // word w restricted to entries in [lo, hi)
    uint32_t m = map[w];
    const int b = w << 5;
    if (lo > b) m &= lo - b >= 32 ? 0u : (0xffffffffu << (lo - b));
    if (hi < b + 32) m &= hi <= b ? 0u : (0xffffffffu >> (b + 32 - hi));
    return m;
}

struct FlatParams {
    const int2* __restrict__ batch;      // nbatch + 1: (first read, histogram words in use) of every batch
    const int* __restrict__ rbase;       // per read: first word of its profile inside its batch, -1 = fallback
    const int* __restrict__ self_cnt;    // ingest: records with A == B per read
    uint32_t* __restrict__ prof;         // scanned packed profiles, kFlatBins words per batch
    int* __restrict__ cov_maxbin;        // last bin of the cut-off-free profile (-1 = empty pile-up)
    int* __restrict__ mean_cov;          // per-read mean coverage, -1 = not part of the estimate
    uint8_t* __restrict__ rflags;
    int* __restrict__ counters1;         // [0] length of big_list (phase 1)
    int* __restrict__ big_list;
    const int* __restrict__ scal;        // [1] = MIN_COV (K2 only)
    int r_begin, r_end;                  // first / last A-read with records
    int v2;                              // K1 ran in its second form (k_profile_flat2)
};

// Which 4-record group of a 1024-record tile a thread takes.  With the identity map the 32
// lanes of a warp hold 128 consecutive records, i.e. one or two reads, and the ~16 of them that
// start in the read's first bin (or end in its last) serialise on one shared-memory word.
// Spreading the lanes over SPREAD windows 128 records apart (32 / SPREAD lanes, a 16 * 32 / SPREAD
// byte run, per window) divides that multiplicity by SPREAD while every window still reads whole
// sectors.  Measured (ms): 1 -> 0.404, 4 -> 0.365, 8 -> 0.368, 16 -> 0.395, 32 -> 0.459.
template <int SPREAD>
__device__ __forceinline__ int flat_group(int tid) {
    if (SPREAD <= 1) return tid;
    constexpr int L = 32 / SPREAD;  // lanes per window
    const int lane = tid & 31, warp = tid >> 5;
    return (lane % L) + L * warp + (kFlatThreads / SPREAD) * (lane / L);
}

// filter.cpp:642-656 (mean over reads >= 5000 bp that have a pile-up) and filter.cpp:552-561
// (self-match reads; float accumulation in record order), shared by K1 and its fallback.
__device__ __forceinline__ void finalize_read(const RecView& rv, const ReadView& rd, const FlatParams& F, int read,
                                              long long sum, int maxbin) {
    const int len0 = maxbin + 1;
    const int mean = (int)(sum / (long long)max(1, len0));
    const int rl = rd.rlen[read];
    F.cov_maxbin[read] = maxbin;
    F.mean_cov[read] = (rl >= 5000 && read >= F.r_begin && read <= F.r_end) ? mean : -1;
    uint8_t f = 0;
    if (self_count(F.self_cnt[read]) > 0) {
        float cov = 0.0f;
        for (int64_t k = rv.read_off[read]; k < rv.read_off[read + 1]; k++) {
            if (rv.bread[k] != read) continue;
            cov = __fadd_rn(cov, (float)(rv.aepos[k] - rv.abpos[k]));
            // B span is strand-invariant: (blen-bbpos) - (blen-bepos) = bepos - bbpos
            cov = __fadd_rn(cov, (float)(rv.bepos[k] - rv.bbpos[k]));
        }
        cov = __fdiv_rn(cov, (float)rl);
        if ((double)cov > 4.5 && rl > 10000) f |= kFlagSelf;
    }
    F.rflags[read] = f;
}

// ------------------------------------------------------------------ K1

template <int SPREAD>
__global__ void __launch_bounds__(kFlatThreads)
k_profile_flat(RecView rv, ReadView rd, hg_filter_params P, FlatParams F) {
    __shared__ __align__(16) uint32_t hist[kFlatWords];
    __shared__ uint32_t wtot[2][kFlatThreads / 32];
    // per read of the batch: sum_records (bin(aepos) - bin(abpos)) = sum_j cov0[j], max bin(aepos) =
    // profile length - 1, and the same maximum without the A == B records for the reads that have some
    __shared__ int sh_sum[kFlatMaxReads], sh_max[kFlatMaxReads], sh_max2[kFlatMaxReads];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int2 bt = F.batch[blockIdx.x];
    const int f0 = bt.x, f1 = F.batch[blockIdx.x + 1].x;
    const int nb = bt.y;
    const int npass = (nb + kFlatPass - 1) / kFlatPass;

    // ---- the batch's records: four per thread and step
    const int64_t k_begin = nb > 0 ? rv.read_off[f0] : 0, k_end = nb > 0 ? rv.read_off[f1] : 0;
    const int64_t g0 = k_begin & ~(int64_t)3;
    const int grp = flat_group<SPREAD>(tid) * 4;
    int4 va = make_int4(-1, -1, -1, -1), vs = make_int4(0, 0, 0, 0), ve = vs;
    auto load4 = [&](int64_t k) {
        if (k >= k_begin && k + 4 <= k_end) {
            va = __ldg(reinterpret_cast<const int4*>(rv.aread + k));
            vs = __ldg(reinterpret_cast<const int4*>(rv.abpos + k));
            ve = __ldg(reinterpret_cast<const int4*>(rv.aepos + k));
        } else {  // ragged ends of the batch: records outside it get read id -1
            int a[4], s[4], e[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int64_t ki = k + i;
                const bool in = ki >= k_begin && ki < k_end;
                a[i] = in ? __ldg(rv.aread + ki) : -1;
                s[i] = in ? __ldg(rv.abpos + ki) : 0;
                e[i] = in ? __ldg(rv.aepos + ki) : 0;
            }
            va = make_int4(a[0], a[1], a[2], a[3]);
            vs = make_int4(s[0], s[1], s[2], s[3]);
            ve = make_int4(e[0], e[1], e[2], e[3]);
        }
    };
    if (g0 < k_end) load4(g0 + grp);  // in flight while the histogram is cleared

    // ---- zero
    for (int j = tid * 4; j < npass * kFlatPass + 16 && j < kFlatWords; j += kFlatThreads * 4)
        sts128(hist + j, make_uint4(0, 0, 0, 0));
    for (int r = tid; r < f1 - f0; r += kFlatThreads) {
        sh_sum[r] = 0;
        sh_max[r] = -1;
    }
    __syncthreads();

    // ---- scatter (profileCoverage, LAInterface.cpp:4298-4320): the next step's loads are
    // issued before this step's events go out.  Every record is scattered, A == B ones included
    // (telling them apart would need the bread column): they are taken out again below.
    const int C = P.cut_off;
    for (int64_t kb = g0; kb < k_end; kb += kFlatThreads * 4) {
        const int a[4] = {va.x, va.y, va.z, va.w}, s[4] = {vs.x, vs.y, vs.z, vs.w};
        const int e[4] = {ve.x, ve.y, ve.z, ve.w};
        if (kb + kFlatThreads * 4 < k_end) load4(kb + kFlatThreads * 4 + grp);
        // records are sorted by A-read: in most groups one lookup of the read's base serves all four
        const int base0 = a[0] >= 0 ? __ldg(F.rbase + a[0]) : -1;
        int cur = -1, acc = 0, mx = -1;  // run of records of one read inside this group
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int base = base0;
            if (a[i] != a[0]) base = a[i] >= 0 ? __ldg(F.rbase + a[i]) : -1;
            if (base < 0) continue;
            // 0 <= abpos < aepos (ingest check)
            const int q_s = s[i] / kReso + 1, q_e = e[i] / kReso + 1;
            const int b_sc = base + cov_bin(s[i] + C, kReso), b_ec = base + cov_bin(e[i] - C, kReso);
            atomicAdd(&hist[sw(base + q_s)], 1u);
            atomicAdd(&hist[sw(base + q_e)], 0u - 1u);
            atomicAdd(&hist[sw(b_sc)], 1u << 16);
            atomicAdd(&hist[sw(b_ec)], 0u - (1u << 16));
            if (a[i] != cur) {
                if (cur >= 0) {
                    atomicAdd(&sh_sum[cur - f0], acc);
                    atomicMax(&sh_max[cur - f0], mx);
                }
                cur = a[i];
                acc = 0;
                mx = -1;
            }
            acc += q_e - q_s;
            mx = max(mx, q_e);
        }
        if (cur >= 0) {
            atomicAdd(&sh_sum[cur - f0], acc);
            atomicMax(&sh_max[cur - f0], mx);
        }
    }
    // A == B records are inactive (filter.cpp:538-547) and rare; their per-read count comes from
    // the ingest.  One thread per such read removes their events again (atomic adds commute, so
    // this needs no barrier) and recomputes the maximum without them.
    for (int r = tid; r < f1 - f0; r += kFlatThreads) {
        const int read = f0 + r;
        const int base = F.rbase[read];
        if (self_count(F.self_cnt[read]) <= 0 || base < 0) continue;
        int mx = -1, acc = 0;
        for (int64_t k = rv.read_off[read]; k < rv.read_off[read + 1]; k++) {
            const int as = rv.abpos[k], ae = rv.aepos[k];
            const int q_s = as / kReso + 1, q_e = ae / kReso + 1;
            if (rv.bread[k] != read) {
                mx = max(mx, q_e);
                continue;
            }
            atomicAdd(&hist[sw(base + q_s)], 0u - 1u);
            atomicAdd(&hist[sw(base + q_e)], 1u);
            atomicAdd(&hist[sw(base + cov_bin(as + C, kReso))], 0u - (1u << 16));
            atomicAdd(&hist[sw(base + cov_bin(ae - C, kReso))], 1u << 16);
            acc += q_e - q_s;
        }
        atomicAdd(&sh_sum[r], -acc);
        sh_max2[r] = mx;
    }
    __syncthreads();

    // ---- one prefix sum over the whole batch; the scanned words also go to HBM for K2
    uint32_t* const pw = F.prof + (size_t)blockIdx.x * kFlatBins;
    uint32_t carry = 0;
    for (int pass = 0; pass < npass; pass++) {
        const int j0 = pass * kFlatPass + tid * kFlatItems;
        const int sx = sw(j0) ^ j0;  // the flipped bits: common to the whole 16-word chunk
        uint32_t v[kFlatItems];
        if (j0 < nb) {
#pragma unroll
            for (int q = 0; q < kFlatItems / 4; q++) {
                const uint4 x = lds128(hist + ((j0 + 4 * q) ^ sx));
                v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
            }
#pragma unroll
            for (int i = 1; i < kFlatItems; i++) v[i] += v[i - 1];
        } else {
#pragma unroll
            for (int i = 0; i < kFlatItems; i++) v[i] = 0;
        }
        const uint32_t incl = warp_incl_scan(v[kFlatItems - 1]);
        if (lane == 31) wtot[pass & 1][warp] = incl;
        __syncthreads();
        uint32_t pre = carry + incl - v[kFlatItems - 1];
#pragma unroll
        for (int w = 0; w < kFlatThreads / 32; w++) {
            const uint32_t t = wtot[pass & 1][w];
            if (w < warp) pre += t;
            carry += t;
        }
        if (j0 < nb) {
#pragma unroll
            for (int i = 0; i < kFlatItems; i++) v[i] += pre;
#pragma unroll
            for (int q = 0; q < kFlatItems / 4; q++) {
                const uint4 x = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                sts128(hist + ((j0 + 4 * q) ^ sx), x);
                *reinterpret_cast<uint4*>(pw + j0 + 4 * q) = x;
            }
        }
    }
    __syncthreads();

    // ---- per read: length and mean of the cut-off-free profile (filter.cpp:642-656)
    for (int read = f0 + tid; read < f1; read += kFlatThreads) {
        const int base = F.rbase[read];
        const int64_t nrec = rv.read_off[read + 1] - rv.read_off[read];
        if (base < 0 || nrec > Packed<uint32_t>::kMaxCount) {
            F.big_list[atomicAdd(&F.counters1[0], 1)] = read;
            continue;
        }
        finalize_read(rv, rd, F, read, sh_sum[read - f0],
                      self_count(F.self_cnt[read]) > 0 ? sh_max2[read - f0] : sh_max[read - f0]);
    }
}

// ------------------------------------------------------------------ K1, second form
//
// The first form above is bound by its shared-memory atomics (ncu, round 1: 10.8 M ATOMS warp
// instructions = 47 M of the 58 M shared-memory wavefronts, l1tex 89 %, HBM 36 %): four per
// record for the two profiles plus the per-read sum / maximum.  This form needs TWO per record
// and none per read.
//
// Both profiles are differences of the same two counting functions (LAInterface.cpp:4298-4320
// counts, for entry j, the events at positions < 40 j):
//     cov0[j] = #{abpos < 40 j}     - #{aepos < 40 j}
//     covC[j] = #{abpos < 40 j - C} - #{aepos < 40 j + C}
// so with C = 300 = 15 x 20 ONE histogram of the record starts and ends on a 20-bp grid (starts in
// the low half of a word, ends in the high half) and ONE exclusive prefix sum P over it give
//     cov0[j] = P_S(2 j) - P_E(2 j),      covC[j] = P_S(2 j - 15) - P_E(2 j + 15).
// The prefix sum runs over the whole batch; a read's own counts are differences of P inside its
// word range (everything before it has both started and ended), taken mod 2^16.
//
//   scatter   two ATOMS per record into hist[2 base + pos / 20]
//   scan      each thread owns 32 histogram words = 16 bins q: E[q] = P(2 q) and Q[q] = P(2 q + 1)
//             stay in registers; covC needs Q[q - 8] and Q[q + 7], which live in the neighbouring
//             threads: 15 shuffles (+ a few words through shared memory at the warp seams)
//   output    the packed (cov0, covC) words K2 expects go to HBM, 64 B per thread
//   per read  a second block-wide scan, over cov0, leaves its running sum T in shared memory;
//             one THREAD per read gets the profile sum as T[end] - T[begin] and the profile length
//             by bisection for the point where T stops growing (filter.cpp:642-656) -- no atomics
//
// No read boundaries are needed in the flat phases: the last 9 words of every read's range
// carry no events (bins_needed), so the look-behind by 8 of a read's first words sees exactly the
// counts at its own start; the look-ahead by 7 of its last words may see the next read's earliest
// ends, which only pushes covC further below zero in bins where it already is <= 0 and where
// nothing but "not above MIN_COV" is ever asked of it (K2 runs this path for MIN_COV >= 0 only).
//
// The profile length is 1 + the last bin with cov0 > 0, which holds unless a record lies inside
// one 40-bp bin (abpos / 40 == aepos / 40); the ingest flags reads that have such a record
// (kSelfDegenerate) and their sum / length come from the per-read fallback.  Batches with more
// than 32767 records (the 16-bit counts could wrap) are left to the fallbacks entirely, here and
// in K2.
constexpr int kV2CutOff = 300;
constexpr int kV2Words = 2 * kFlatBins + 64;
constexpr int kV2MaxBatchRecords = 32767;

// Histogram layout: a thread of the scan owns 32 consecutive words = 8 vectors, so the 8 lanes
// of a quarter warp are 128 B apart; XOR-ing the vector index with the thread index spreads them
// over the eight 16-byte bank groups.
__device__ __forceinline__ int swh(int j) { return j ^ (((j >> 5) & 7) << 2); }

template <int SPREAD, int MINBLOCKS>
__global__ void __launch_bounds__(kFlatThreads, MINBLOCKS)
k_profile_flat2(RecView rv, ReadView rd, hg_filter_params P, FlatParams F) {
    __shared__ __align__(16) uint32_t buf[kV2Words];
    __shared__ uint32_t wtot[2][kFlatThreads / 32];
    __shared__ uint32_t seam_hi[kFlatThreads / 32][8], seam_lo[kFlatThreads / 32][8];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kFlatThreads / 32;
    const int2 bt = F.batch[blockIdx.x];
    const int f0 = bt.x, f1 = F.batch[blockIdx.x + 1].x;
    const int64_t k_begin = bt.y > 0 ? rv.read_off[f0] : 0, k_end = bt.y > 0 ? rv.read_off[f1] : 0;
    const int nb = (k_end - k_begin > kV2MaxBatchRecords) ? 0 : bt.y;

    if (nb > 0) {
        // ---- the batch's records: four per thread and step (abpos / aepos; aread for the base)
        const int64_t g0 = k_begin & ~(int64_t)3;
        const int grp = flat_group<SPREAD>(tid) * 4;
        int4 va = make_int4(-1, -1, -1, -1), vs = make_int4(0, 0, 0, 0), ve = vs;
        auto load4 = [&](int64_t k) {
            if (k >= k_begin && k + 4 <= k_end) {
                va = __ldg(reinterpret_cast<const int4*>(rv.aread + k));
                vs = __ldg(reinterpret_cast<const int4*>(rv.abpos + k));
                ve = __ldg(reinterpret_cast<const int4*>(rv.aepos + k));
            } else {  // ragged ends of the batch: records outside it get read id -1
                int a[4], s[4], e[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int64_t ki = k + i;
                    const bool in = ki >= k_begin && ki < k_end;
                    a[i] = in ? __ldg(rv.aread + ki) : -1;
                    s[i] = in ? __ldg(rv.abpos + ki) : 0;
                    e[i] = in ? __ldg(rv.aepos + ki) : 0;
                }
                va = make_int4(a[0], a[1], a[2], a[3]);
                vs = make_int4(s[0], s[1], s[2], s[3]);
                ve = make_int4(e[0], e[1], e[2], e[3]);
            }
        };
        if (g0 < k_end) load4(g0 + grp);  // in flight while the histogram is cleared

        // ---- zero (swh permutes inside aligned 256-word blocks)
        const int nzero = min((2 * nb + 255) & ~255, 2 * kFlatBins);
        for (int j = tid * 4; j < nzero; j += kFlatThreads * 4) sts128(buf + j, make_uint4(0, 0, 0, 0));
        __syncthreads();

        // ---- scatter: start -> low half, end -> high half of hist[2 base + pos / 20]
        for (int64_t kb = g0; kb < k_end; kb += kFlatThreads * 4) {
            const int a[4] = {va.x, va.y, va.z, va.w}, s[4] = {vs.x, vs.y, vs.z, vs.w};
            const int e[4] = {ve.x, ve.y, ve.z, ve.w};
            if (kb + kFlatThreads * 4 < k_end) load4(kb + kFlatThreads * 4 + grp);
            // records are sorted by A-read: in most groups one lookup of the read's base serves all four
            const int base0 = a[0] >= 0 ? __ldg(F.rbase + a[0]) : -1;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                int base = base0;
                if (a[i] != a[0]) base = a[i] >= 0 ? __ldg(F.rbase + a[i]) : -1;
                if (base < 0) continue;
                // 0 <= abpos < aepos <= rlen (ingest check)
                atomicAdd(&buf[swh(2 * base + (int)((unsigned)s[i] / 20u))], 1u);
                atomicAdd(&buf[swh(2 * base + (int)((unsigned)e[i] / 20u))], 1u << 16);
            }
        }
        // A == B records are inactive (filter.cpp:538-547) and rare; their per-read count comes
        // from the ingest.  One thread per such read takes their events out again.
        for (int r = tid; r < f1 - f0; r += kFlatThreads) {
            const int read = f0 + r;
            const int base = F.rbase[read];
            if (self_count(F.self_cnt[read]) <= 0 || base < 0) continue;
            for (int64_t k = rv.read_off[read]; k < rv.read_off[read + 1]; k++) {
                if (rv.bread[k] != read) continue;
                atomicAdd(&buf[swh(2 * base + (int)((unsigned)rv.abpos[k] / 20u))], 0u - 1u);
                atomicAdd(&buf[swh(2 * base + (int)((unsigned)rv.aepos[k] / 20u))], 0u - (1u << 16));
            }
        }
        __syncthreads();

        // ---- exclusive prefix sum over the 20-bp histogram, one pass: thread t owns bins
        // [16 t, 16 t + 16) = words [32 t, 32 t + 32).  Threads past the batch's last bin carry the
        // total (their words count as empty): the look-ahead of the last bins reads them.
        uint32_t he[kFlatItems], ho[kFlatItems];
        const int q0 = tid * kFlatItems;
        const bool mine = q0 < nb;
        uint32_t tot = 0;
        if (mine) {
#pragma unroll
            for (int v = 0; v < 8; v++) {
                const uint4 x = lds128(buf + 32 * tid + 4 * (v ^ (tid & 7)));
                he[2 * v] = x.x; ho[2 * v] = x.y; he[2 * v + 1] = x.z; ho[2 * v + 1] = x.w;
            }
#pragma unroll
            for (int i = 0; i < kFlatItems; i++) {
                const uint32_t e0 = he[i], o0 = ho[i];
                he[i] = tot;        // E[q] = P(2 q)
                ho[i] = tot + e0;   // Q[q] = P(2 q + 1)
                tot += e0 + o0;
            }
        } else {
#pragma unroll
            for (int i = 0; i < kFlatItems; i++) he[i] = ho[i] = 0;
        }
        uint32_t batch_total = 0;  // P at the end of the batch: what a look-ahead past the last word sees
        {
            const uint32_t incl = warp_incl_scan(tot);
            if (lane == 31) wtot[0][warp] = incl;
            __syncthreads();  // every histogram word has been read: the buffer is free for T
            uint32_t pre = incl - tot;
#pragma unroll
            for (int w = 0; w < NW; w++) {
                const uint32_t t = wtot[0][w];
                if (w < warp) pre += t;
                batch_total += t;
            }
#pragma unroll
            for (int i = 0; i < kFlatItems; i++) {
                he[i] += pre;
                ho[i] += pre;
            }
        }
        // the Q values the neighbouring warps need: Q[q0 + 8 .. q0 + 15] of a warp's last lane for
        // the look-behind of the next warp's first lane, Q[q0 .. q0 + 6] of its first lane for the
        // look-ahead of the previous warp's last lane
        if (lane == 31) {
#pragma unroll
            for (int i = 0; i < 8; i++) seam_hi[warp][i] = ho[8 + i];
        }
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 7; i++) seam_lo[warp][i] = ho[i];
        }
        __syncthreads();

        // ---- the packed words: cov0[q] = E_S[q] - E_E[q], covC[q] = Q_S[q - 8] - Q_E[q + 7]
        uint32_t word[kFlatItems];
        uint32_t csum = 0;
#pragma unroll
        for (int i = 0; i < kFlatItems; i++) {
            uint32_t ps, pe;
            if (i < 8) {  // look-behind into the previous thread
                ps = __shfl_up_sync(0xffffffffu, ho[8 + i], 1);
                if (lane == 0) ps = warp > 0 ? seam_hi[warp - 1][i] : 0u;
            } else {
                ps = ho[i - 8];
            }
            if (i + 7 < kFlatItems) {
                pe = ho[i + 7];
            } else {  // look-ahead into the next thread
                pe = __shfl_down_sync(0xffffffffu, ho[i + 7 - kFlatItems], 1);
                if (lane == 31) pe = warp + 1 < NW ? seam_lo[warp + 1][i + 7 - kFlatItems] : batch_total;
            }
            const uint32_t c0 = (he[i] - (he[i] >> 16)) & 0xffffu;
            const uint32_t c1 = (ps - (pe >> 16)) & 0xffffu;
            word[i] = c0 | (c1 << 16);
            he[i] = csum;  // from here on: the thread-local exclusive prefix of cov0
            csum += c0;
        }
        uint32_t* const pw = F.prof + (size_t)blockIdx.x * kFlatBins;
        if (mine) {
#pragma unroll
            for (int v = 0; v < kFlatItems / 4; v++)
                *reinterpret_cast<uint4*>(pw + q0 + 4 * v) =
                    make_uint4(word[4 * v], word[4 * v + 1], word[4 * v + 2], word[4 * v + 3]);
        }

        // ---- T = exclusive prefix sum of cov0 over the batch, into shared memory (swizzled like the
        // first form's histogram: 16 words per thread)
        {
            const uint32_t incl = warp_incl_scan(csum);
            if (lane == 31) wtot[1][warp] = incl;
            __syncthreads();
            uint32_t pre = incl - csum, total = 0;
#pragma unroll
            for (int w = 0; w < NW; w++) {
                const uint32_t t = wtot[1][w];
                if (w < warp) pre += t;
                total += t;
            }
            if (mine) {
                const int sx = sw(q0) ^ q0;  // the flipped bits: common to the whole 16-word chunk
#pragma unroll
                for (int v = 0; v < kFlatItems / 4; v++)
                    sts128(buf + ((q0 + 4 * v) ^ sx), make_uint4(he[4 * v] + pre, he[4 * v + 1] + pre,
                                                                 he[4 * v + 2] + pre, he[4 * v + 3] + pre));
            }
            if (tid == 0 && (nb & (kFlatItems - 1)) == 0) buf[sw(nb)] = total;  // T[nb]: its chunk was skipped
        }
        __syncthreads();
    }

    // ---- per read, one thread each: sum and length of the cut-off-free profile (filter.cpp:642-656)
    for (int read = f0 + tid; read < f1; read += kFlatThreads) {
        const int base = F.rbase[read];
        const int64_t nrec = rv.read_off[read + 1] - rv.read_off[read];
        if (base < 0 || nb == 0 || nrec > Packed<uint32_t>::kMaxCount || is_degenerate(F.self_cnt[read])) {
            // too long / too deep / a record inside one bin: sum and length from the records
            F.big_list[atomicAdd(&F.counters1[0], 1)] = read;
            continue;
        }
        const int nbz = bins_needed(rd.rlen[read], P);
        const uint32_t t_begin = buf[sw(base)], t_end = buf[sw(base + nbz)];
        // smallest j with T[base + j] == t_end: bins j - 1 is the last one with cov0 > 0
        int lo = 0, hi = nbz;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (buf[sw(base + mid)] == t_end) hi = mid; else lo = mid + 1;
        }
        finalize_read(rv, rd, F, read, (long long)(t_end - t_begin), lo == 0 ? -1 : lo);
    }
}

// Fallback of K1 for the reads the flat path cannot take: one warp per read, the same sums
// straight from the records:  sum_j cov[j] = sum_records (bin(aepos) - bin(abpos)),
// length = max bin(aepos) + 1.
__global__ void __launch_bounds__(128)
k_cov_big(RecView rv, ReadView rd, FlatParams F) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int nbig = F.counters1[0];
    for (int w = warp; w < nbig; w += nwarps) {
        const int read = F.big_list[w];
        long long sum = 0;
        int mx = -1;
        for (int64_t k = rv.read_off[read] + lane_id(); k < rv.read_off[read + 1]; k += 32) {
            if (rv.bread[k] == read) continue;
            const int be = cov_bin(rv.aepos[k], kReso);
            sum += be - cov_bin(rv.abpos[k], kReso);
            mx = max(mx, be);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        mx = warp_max(mx);
        if (lane_id() == 0) finalize_read(rv, rd, F, read, sum, mx);
    }
}

// ------------------------------------------------------------------ K2

template <bool DUMP>
__global__ void __launch_bounds__(kFlatThreads, 8)
k_mask_anno_flat(RecView rv, ReadView rd, hg_filter_params P, FlatParams F, MaskAnnoOut out) {
    __shared__ __align__(16) uint16_t zmap16[kFlatMaps];  // bin has cut-off coverage <= MIN_COV
    __shared__ __align__(16) uint16_t cmap16[kFlatMaps];  // |cov0[j] - cov0[j-1]| above the smallest threshold
    const uint32_t* zmap = reinterpret_cast<const uint32_t*>(zmap16);
    const uint32_t* cmap = reinterpret_cast<const uint32_t*>(cmap16);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int2 bt = F.batch[blockIdx.x];
    const int f0 = bt.x, f1 = F.batch[blockIdx.x + 1].x;
    const int MIN_COV = F.scal[1];
    constexpr int reso = kReso;
    // MIN_COV < 0: runs could cross read boundaries, everything goes the generic way; so do the
    // batches the second form of K1 declined (more records than its 16-bit prefix counts hold)
    const bool declined = F.v2 && bt.y > 0 && rv.read_off[f1] - rv.read_off[f0] > kV2MaxBatchRecords;
    const int nb = (MIN_COV < 0 || declined) ? 0 : bt.y;
    const int npass = (nb + kFlatPass - 1) / kFlatPass;
    const uint32_t* __restrict__ const pw = F.prof + (size_t)blockIdx.x * kFlatBins;

    // ---- the two bit maps, flat over the batch's bins.  zero <=> high half <= MIN_COV <=> the
    // word, as a signed integer, is below (MIN_COV + 1) << 16 (the low half is >= 0).
    const int zthr = (MIN_COV + 1 > 32767 ? 32767 : MIN_COV + 1) << 16;
    const int RJ = min(P.min_repeat_annotation_threshold, P.max_repeat_annotation_threshold);
    for (int pass = 0; pass < npass; pass++) {
        const int j0 = pass * kFlatPass + tid * kFlatItems;
        uint32_t zbits = 0, cbits = 0;
        uint32_t v[kFlatItems];
        if (j0 < nb) {
#pragma unroll
            for (int q = 0; q < kFlatItems / 4; q++) {
                const uint4 x = __ldg(reinterpret_cast<const uint4*>(pw + j0 + 4 * q));
                v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < kFlatItems; i++) v[i] = 0;
        }
        // the word before this thread's chunk (chunks of one warp are contiguous)
        uint32_t prev = __shfl_up_sync(0xffffffffu, v[kFlatItems - 1], 1);
        if (lane == 0) prev = (j0 > 0 && j0 < nb) ? __ldg(pw + j0 - 1) : 0u;
        if (j0 < nb) {
#pragma unroll
            for (int i = 0; i < kFlatItems; i++) {
                zbits |= ((int)v[i] < zthr) ? (1u << i) : 0u;
                const int g = f_lo(v[i]) - f_lo(i ? v[i - 1] : prev);  // cov0[j0 + i] - cov0[j0 + i - 1]
                cbits |= (g > RJ || g < -RJ) ? (1u << i) : 0u;
            }
        }
        zmap16[pass * kFlatThreads + tid] = (uint16_t)zbits;
        cmap16[pass * kFlatThreads + tid] = (uint16_t)cbits;
    }
    __syncthreads();

    // ---- per read, one thread each.  A read's walk is a chain of dependent steps (bit-map words,
    // coverage look-ups behind the annotation candidates), so the reads of a batch are dealt out
    // round-robin over the CTA's WARPS: consecutive thread ids would put a long-read batch's six reads
    // into one warp -- one instruction stream per CTA, eight per SM -- where this gives six.  (A
    // warp-wide form of the walk, lanes over bit-map words and candidate bits, was tried on the
    // long-read set and lost: 0.95 ms against 0.61 ms; the walk is short on parallel work and the
    // reductions cost more issue slots than they save.)
    const int NHR = P.no_hinge_region;
    const int MINT = P.min_repeat_annotation_threshold, MAXT = P.max_repeat_annotation_threshold;
    constexpr int kWarps = kFlatThreads / 32;
    for (int slot = lane * kWarps + warp; slot < f1 - f0; slot += kFlatThreads) {
        const int read = f0 + slot;
        const int base = F.rbase[read];
        const int64_t nrec = rv.read_off[read + 1] - rv.read_off[read];
        if (base < 0 || nb == 0 || nrec > Packed<uint32_t>::kMaxCount) {
            out.big_list[atomicAdd(&out.counters[3], 1)] = read;
            continue;
        }
        const int nbz = bins_needed(rd.rlen[read], P);
        const int L0 = F.cov_maxbin[read] + 1;  // length of the cut-off-free profile
        auto H = [&](int j) { return __ldg(pw + base + j); };  // packed coverage of the read's bin j (L1 hit)

        // longest run of covered bins (filter.cpp:696-728): the run between two consecutive zeros
        // p < z scores 40 (z - p - 2); bin 0 acts as a zero; '>' keeps the earliest of the longest
        int p = base, bestgap = 0, bestz = 0;
        const int end = base + nbz;
        for (int w = base >> 5; w <= (end - 1) >> 5; w++) {
            uint32_t m = map_word(zmap, w, base + 1, end);
            while (m) {
                const int bit = __ffs(m) - 1;
                const int z = (w << 5) + bit;
                if (z - p > bestgap) {
                    bestgap = z - p;
                    bestz = z;
                }
                // the zeros that follow z back to back each have gap 1: skip them in one go
                const uint32_t t = ~(m >> bit);                     // bit 0 clear
                const int run = t ? __ffs(t) - 1 : 32 - bit;        // consecutive zeros from z on
                p = z + run - 1;
                m = bit + run >= 32 ? 0u : (m >> (bit + run)) << (bit + run);
            }
        }
        int maxstart = 0, maxend = 0, msc = 0, mec = 0;
        if (bestgap >= 3) {
            const int z = bestz - base, pz = z - bestgap;
            msc = pz + 1;
            mec = z - 1;
            maxstart = reso * (pz + 1);
            maxend = reso * (z - 1);
        }

        // telomere / coverage-imbalance flag (filter.cpp:731-760)
        uint8_t flags = 0;
        if (P.delete_telomere) {
            flags = out.rflags[read] & kFlagSelf;
            int limit, div;
            if (mec - msc + 1 > 20) {
                limit = 10;
                div = 10;
            } else {
                limit = (mec - msc) / 2;
                div = limit;
            }
            int sc = 0, ec = 0;
            for (int t = 0; t < limit; t++) {
                sc += max(f_hi(H(msc + t)), MIN_COV);
                ec += max(f_hi(H(mec - t)), MIN_COV);
            }
            if (div == 0) {
                sc = 0;
                ec = 0;
            } else {
                sc /= div;
                ec /= div;
            }
            if (sc >= 10 * ec || ec >= 10 * sc) flags |= kFlagCov;
        }

        // final mask (filter.cpp:777-788)
        const int2 q = rd.qvmask[read];
        int2 mk;
        if (P.use_qv_mask && P.use_coverage_mask)
            mk = make_int2(max(maxstart, q.x), min(maxend, q.y));
        else if (P.use_coverage_mask && !P.use_qv_mask)
            mk = make_int2(maxstart, maxend);
        else
            mk = q;

        // repeat annotation from the coverage gradient (filter.cpp:796-813) + merge pass
        // (filter.cpp:817-829) as a stream: first count what survives, then write it.
        // Candidates are bins j < L0 - 2 with 40 j in [mask.start + NHR, mask.end - NHR];
        // map entry j + 1 flags the jump cov0[j + 1] - cov0[j].
        const int ja_lo = mk.x + NHR <= 0 ? 0 : (mk.x + NHR + reso - 1) / reso;
        const int ja_hi = mk.y - NHR < 0 ? -1 : min((mk.y - NHR) / reso, L0 - 3);
        const int GAP = P.repeat_annotation_gap_threshold;
        int kept = 0, off = 0;
        for (int wr = 0; wr < 2; wr++) {
            int n = 0;
            unsigned cur = 0;
            bool have = false;
            if (ja_hi >= ja_lo) {
                const int lo = base + ja_lo + 1, hi = base + ja_hi + 2;
                for (int w = lo >> 5; w <= (hi - 1) >> 5; w++) {
                    uint32_t m = map_word(cmap, w, lo, hi);
                    while (m) {
                        const int bit = __ffs(m) - 1;
                        m &= m - 1;
                        const int j = (w << 5) + bit - 1 - base;
                        const int c0 = f_lo(H(j));
                        const int g = f_lo(H(j + 1)) - c0;
                        const int thr = min(max((c0 + MIN_COV) / P.coverage_fraction, MINT), MAXT);
                        const int type = g > thr ? 1 : (g < -thr ? -1 : 0);
                        if (type == 0) continue;
                        const unsigned nxt = ((unsigned)(reso * j) << 2) | (unsigned)(type + 1);
                        if (!have) {
                            cur = nxt;
                            have = true;
                            continue;
                        }
                        const int ct = (int)(cur & 3u) - 1;
                        const int gap = (int)(nxt >> 2) - (int)(cur >> 2);
                        if (ct == 1 && type == 1 && gap < GAP) {
                            continue;   // +1,+1 close together: the later one goes
                        } else if (ct == -1 && type == -1 && gap < GAP) {
                            cur = nxt;  // -1,-1 close together: the earlier one goes
                        } else {
                            if (wr) {
                                out.anno_pool[off + n] = make_int2((int)(cur >> 2), (int)(cur & 3u) - 1);
                                out.hinge_keep[off + n] = 0;
                            }
                            n++;
                            cur = nxt;
                        }
                    }
                }
            }
            if (have) {
                if (wr) {
                    out.anno_pool[off + n] = make_int2((int)(cur >> 2), (int)(cur & 3u) - 1);
                    out.hinge_keep[off + n] = 0;
                }
                n++;
            }
            if (wr == 0) {
                kept = n;
                if (kept == 0) break;
                off = atomicAdd(&out.counters[0], kept);
                if (off + kept > out.anno_cap) {
                    atomicExch(&out.counters[2], 1);
                    off = -1;
                    break;
                }
            }
        }

        // hinge pre-test: mean coverage near both mask ends (filter.cpp:842-865); its outcome
        // only matters for reads that carry annotations
        bool skip_hinges = false;
        if (kept > 0) {
            int cs = 0, ns = 0, ce = 0, ne = 0;
            int jlo = mk.x <= 0 ? 0 : (mk.x + reso - 1) / reso;  // bins with mk.x <= 40 j <= mk.x + NHR
            int jhi = mk.x + NHR < 0 ? -1 : min((mk.x + NHR) / reso, L0 - 1);
            for (int j = jlo; j <= jhi; j++) {
                cs += f_lo(H(j));
                ns++;
            }
            jlo = mk.y - NHR <= 0 ? 0 : (mk.y - NHR + reso - 1) / reso;  // mk.y - NHR <= 40 j <= mk.y
            jhi = mk.y < 0 ? -1 : min(mk.y / reso, L0 - 1);
            for (int j = jlo; j <= jhi; j++) {
                ce += f_lo(H(j));
                ne++;
            }
            // float on purpose: 0/0 = NaN makes the '< 10' test false (filter.cpp:861-865)
            const float avg_end = __fdiv_rn((float)ce, (float)ne);
            const float avg_start = __fdiv_rn((float)cs, (float)ns);
            skip_hinges = fabsf(__fsub_rn(avg_end, avg_start)) < 10.0f;
        }

        store_mask(out, read, mk);
        out.cmask[read] = make_int2(msc, mec);
        out.rflags[read] = flags | (skip_hinges ? kFlagSkipHinge : 0);
        out.anno_ref[read] = make_int2(off, kept);
        if (kept > 0 && !skip_hinges && off >= 0)
            push_work_item(out, read, rv.read_off[read], (int)nrec, mk, off, kept);
    }

    // ---- optional dump of the cut-off-free profiles for .coverage.txt (filter.cpp:599-602)
    if (DUMP && nb > 0) {
        for (int read = f0 + warp; read < f1; read += kFlatThreads / 32) {
            const int base = F.rbase[read];
            if (base < 0 || rv.read_off[read + 1] - rv.read_off[read] > Packed<uint32_t>::kMaxCount) continue;
            const int L0 = F.cov_maxbin[read] + 1;
            int* dst = out.cov0 + out.cov0_off[read];
            for (int j = lane; j < L0; j += 32) dst[j] = f_lo(__ldg(pw + base + j));
        }
    }
}

// ------------------------------------------------------------------ host side

// Greedy packing of the reads [lo, hi) into batches of at most kFlatBins histogram words and
// kFlatMaxReads reads.
//   batch  (first read, words in use) per batch, closed by (hi, 0)
//   rbase  per read: first word of its profile inside its batch; -1 = fallback path
void flat_plan(const int* rlen, int lo, int hi, int n_read, int cut_off, std::vector<int2>* batch,
               std::vector<int>* rbase) {
    batch->clear();
    rbase->assign((size_t)n_read, -1);
    int used = 0;
    for (int r = lo; r < hi; r++) {
        const int nbz = bins_needed(rlen[r], cut_off);
        const bool fits = nbz <= kFlatBins;  // otherwise fallback; still belongs to a batch, which reports it
        if (batch->empty() || (fits && used + nbz > kFlatBins) || r - batch->back().x >= kFlatMaxReads) {
            batch->push_back(make_int2(r, 0));
            used = 0;
        }
        if (!fits) continue;
        (*rbase)[r] = used;
        used += nbz;
        batch->back().y = used;
    }
    if (batch->empty()) batch->push_back(make_int2(lo, 0));
    batch->push_back(make_int2(hi, 0));
}

// The second form of K1 is written for the nominal cut-off (300: every INI the reference ships).
// It halves the shared-memory atomics per record but handles two histogram words per coverage bin,
// so it only pays where records outnumber bins (short reads, deep pile-ups: 0.354 vs 0.347 ms on
// the 62 M-record N(3500,1500) set, a wash); with long reads (3 bins per record on the
// N(24000,8000) set) the first form is faster (0.92 vs 1.14 ms).  flat_bins_per_record is set
// when the batch plan is made.
static bool use_v2(const FilterScratch& s, const hg_filter_params& P) {
    if (s.flat_kernel == 1 || P.cut_off != kV2CutOff) return false;
    if (s.flat_kernel >= 5) return true;
    return s.flat_bins_per_record < 1.5f;
}

static FlatParams flat_params(const FilterScratch& s, const hg_filter_params& P, int r_begin, int r_end) {
    FlatParams F;
    F.v2 = use_v2(s, P) ? 1 : 0;
    F.batch = s.flat_batch;
    F.rbase = s.flat_rbase;
    F.self_cnt = s.self_cnt;
    F.prof = s.flat_prof;
    F.cov_maxbin = s.cov_maxbin;
    F.mean_cov = s.mean_cov;
    F.rflags = s.rflags;
    F.counters1 = s.counters1;
    F.big_list = s.big_list;
    F.scal = s.scal;
    F.r_begin = r_begin;
    F.r_end = r_end;
    return F;
}

// Phase 1 of the stage: both coverage profiles of every owned read, their lengths and means.
void launch_profile(const RecView& rv, const ReadView& rd, const hg_filter_params& P, int r_begin,
                    int r_end, FilterScratch& s, cudaStream_t st) {
    const FlatParams F = flat_params(s, P, r_begin, r_end);
    // the counters of both phases; per-read results of reads outside the planned range were
    // cleared when the plan was made (hg_capi.cu)
    cudaMemsetAsync(s.counters, 0, sizeof(int) * 16, st);
    const int grid = s.flat_nbatch;
    if (grid <= 0) return;
    g_launches += 2;
    if (F.v2) {
        // tuning aids: resident CTAs per SM the compiler aims for (registers), scatter spread
        if (s.flat_kernel == 5) k_profile_flat2<8, 4><<<grid, kFlatThreads, 0, st>>>(rv, rd, P, F);
        else if (s.flat_kernel == 6) k_profile_flat2<8, 6><<<grid, kFlatThreads, 0, st>>>(rv, rd, P, F);
        else if (s.flat_spread == 1) k_profile_flat2<1, 4><<<grid, kFlatThreads, 0, st>>>(rv, rd, P, F);
        else if (s.flat_spread == 4) k_profile_flat2<4, 4><<<grid, kFlatThreads, 0, st>>>(rv, rd, P, F);
        else if (s.flat_spread == 16) k_profile_flat2<16, 4><<<grid, kFlatThreads, 0, st>>>(rv, rd, P, F);
        else k_profile_flat2<8, 5><<<grid, kFlatThreads, 0, st>>>(rv, rd, P, F);
    }
    else switch (s.flat_spread) {
        case 1: k_profile_flat<1><<<grid, kFlatThreads, 0, st>>>(rv, rd, P, F); break;
        case 4: k_profile_flat<4><<<grid, kFlatThreads, 0, st>>>(rv, rd, P, F); break;
        case 16: k_profile_flat<16><<<grid, kFlatThreads, 0, st>>>(rv, rd, P, F); break;
        default: k_profile_flat<8><<<grid, kFlatThreads, 0, st>>>(rv, rd, P, F); break;
    }
    k_cov_big<<<16, 128, 0, st>>>(rv, rd, F);
}

// Phase 2 after the median: masks, annotations, hinge work list (the generic fallback kernel
// for the reads it reports is launched by launch_mask_anno, hg_filter.cu).
void launch_mask_anno_flat(const RecView& rv, const ReadView& rd, const hg_filter_params& P, int r_begin,
                           int r_end, FilterScratch& s, const MaskAnnoOut& out, cudaStream_t st) {
    const FlatParams F = flat_params(s, P, r_begin, r_end);
    const int grid = s.flat_nbatch;
    if (grid <= 0) return;
    g_launches += 1;
    if (out.cov0)
        k_mask_anno_flat<true><<<grid, kFlatThreads, 0, st>>>(rv, rd, P, F, out);
    else
        k_mask_anno_flat<false><<<grid, kFlatThreads, 0, st>>>(rv, rd, P, F, out);
}

}  // namespace hg
