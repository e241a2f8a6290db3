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
//   K2 (after the median), two kernels:
//     k_mask_bits_flat  streams the stored profiles and leaves two bits per bin in HBM: bins whose
//       cut-off coverage is <= MIN_COV ("zeros") and bins where the coverage jumps by more than the
//       smallest annotation threshold
//     k_mask_walk       one thread per read walks its slice of the maps: longest covered run
//       (filter.cpp:696-728), mask, telomere flag, repeat annotations with the streaming form of the
//       merge pass (filter.cpp:796-829), hinge pre-test (filter.cpp:842-865)
//
// K1 is bound by the shared-memory data pipe (ncu: l1tex data-pipe wavefronts > 70 % of peak; ~3
// wavefronts per ATOMS on random bins is what the banks give, same-address lanes of a +1 are merged
// by the hardware (ATOMS.POPC.INC), those of other addends are not), so everything else in it is kept
// conflict-free (sw) and the global accesses wide (128-bit loads, 256-bit stores).  Tried and
// measured on B200 on the way: match.any aggregation of equal bins 0.84 ms, ballot aggregation of the
// hot bins 0.54 ms, none 0.40 ms, + lane spreading 0.365 ms, + swizzle and prefetch 0.34 ms, + 256-bit
// stores 0.32 ms (short-read set).  A third, TMA-staged form of K1 follows the two flat ones.
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
constexpr int kSlabRecords = 4096;                       // records of a batch (third form of K1: staged by TMA)

__device__ __forceinline__ uint4 lds128(const uint32_t* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void sts128(uint32_t* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }
// 256-bit store (sm_100: STG.E.ENL2.256): a thread's eight consecutive profile words fill a whole 32-byte
// sector, where four 128-bit stores 64 B apart leave every sector half written per request
__device__ __forceinline__ void st_global_v8(uint32_t* p, const uint32_t* v) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void ld_global_nc_v8(const uint32_t* p, uint32_t* v) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ int f_lo(uint32_t v) { return (int)(v & 0xffffu); }   // after the scan: cov0 >= 0
__device__ __forceinline__ int f_hi(uint32_t v) { return (int)v >> 16; }

// Shared-memory layout of the histogram.  A thread of the scan owns 16 consecutive words and
// moves them as four 128-bit vectors; with the identity layout the 8 lanes of a quarter warp
// start 64 B apart and hit only two of the eight 16-byte bank groups (ncu: twice the ideal
// wavefronts on every LDS.128 / STS.128).  Flipping word-index bits 2-3 with bits 5-6 keeps
// every aligned 4-word group intact and makes those accesses conflict-free.
__device__ __forceinline__ int sw(int j) { return j ^ ((j >> 3) & 12); }

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
    const int* __restrict__ rbatch;      // per read: its batch
    const int4* __restrict__ desc;       // per batch: (first read, reads, words, 0) (first record lo, hi, records, 0)
    uint16_t* __restrict__ zmap;         // K2 bit maps, kFlatMaps 16-bit entries per batch
    uint16_t* __restrict__ cmap;
    int p_lo, p_hi;                      // planned read range
    int nbatch;
    int r_begin, r_end;                  // first / last A-read with records
    int v2;                              // form K1 ran in: 0 first, 1 second (k_profile_flat2), 2 third (k_profile_tma)
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
__device__ __forceinline__ void finalize_read(const RecView& rv, const FlatParams& F, int read, long long sum,
                                              int maxbin, int rl, int self_cnt) {
    const int len0 = maxbin + 1;
    const int mean = (int)(sum / (long long)max(1, len0));
    F.cov_maxbin[read] = maxbin;
    F.mean_cov[read] = (rl >= 5000 && read >= F.r_begin && read <= F.r_end) ? mean : -1;
    uint8_t f = 0;
    if (self_count(self_cnt) > 0) {
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
__device__ __forceinline__ void finalize_read(const RecView& rv, const ReadView& rd, const FlatParams& F, int read,
                                              long long sum, int maxbin) {
    finalize_read(rv, F, read, sum, maxbin, rd.rlen[read], F.self_cnt[read]);
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
    // full steps spread the lanes over record windows (flat_group); the last, partial step of a batch
    // takes the records in thread order, so that only the warps that still have records run its body
    const int grp_spread = flat_group<SPREAD>(tid) * 4, grp_tail = tid * 4;
    auto grp_of = [&](int64_t kb) { return k_end - kb >= kFlatThreads * 4 ? grp_spread : grp_tail; };
    int4 va = make_int4(-1, -1, -1, -1), vs = make_int4(0, 0, 0, 0), ve = vs;
    auto load4 = [&](int64_t k) {
        if (k >= k_end) {
            va = make_int4(-1, -1, -1, -1);
        } else if (k >= k_begin && k + 4 <= k_end) {
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
    if (g0 < k_end) load4(g0 + grp_of(g0));  // in flight while the histogram is cleared

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
        if (kb + kFlatThreads * 4 < k_end) load4(kb + kFlatThreads * 4 + grp_of(kb + kFlatThreads * 4));
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
            for (int q = 0; q < kFlatItems / 8; q++) st_global_v8(pw + j0 + 8 * q, v + 8 * q);
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
        // (lane mapping of full steps and of the last, partial one: see the first form)
        const int grp_spread = flat_group<SPREAD>(tid) * 4, grp_tail = tid * 4;
        auto grp_of = [&](int64_t kb) { return k_end - kb >= kFlatThreads * 4 ? grp_spread : grp_tail; };
        int4 va = make_int4(-1, -1, -1, -1), vs = make_int4(0, 0, 0, 0), ve = vs;
        auto load4 = [&](int64_t k) {
            if (k >= k_end) {
                va = make_int4(-1, -1, -1, -1);
            } else if (k >= k_begin && k + 4 <= k_end) {
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
        if (g0 < k_end) load4(g0 + grp_of(g0));  // in flight while the histogram is cleared

        // ---- zero (swh permutes inside aligned 256-word blocks)
        const int nzero = min((2 * nb + 255) & ~255, 2 * kFlatBins);
        for (int j = tid * 4; j < nzero; j += kFlatThreads * 4) sts128(buf + j, make_uint4(0, 0, 0, 0));
        __syncthreads();

        // ---- scatter: start -> low half, end -> high half of hist[2 base + pos / 20]
        for (int64_t kb = g0; kb < k_end; kb += kFlatThreads * 4) {
            const int a[4] = {va.x, va.y, va.z, va.w}, s[4] = {vs.x, vs.y, vs.z, vs.w};
            const int e[4] = {ve.x, ve.y, ve.z, ve.w};
            if (kb + kFlatThreads * 4 < k_end) load4(kb + kFlatThreads * 4 + grp_of(kb + kFlatThreads * 4));
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
            for (int v = 0; v < kFlatItems / 8; v++) st_global_v8(pw + q0 + 8 * v, word + 8 * v);
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

// ------------------------------------------------------------------ K1, third form
//
// The first two forms wait on their own global loads (ncu, round 2: long-scoreboard is the top
// stall, 1.15 eligible warps per scheduler, issue slots 53 % busy).  This form is a PERSISTENT CTA
// that never waits on a global load in its steady state and never looks at `aread`:
//
//   staging   the batch's abpos / aepos columns (one contiguous run of <= kSlabRecords records, the
//             plan sees to that) and its per-read tables (CSR offsets, histogram bases, read lengths)
//             arrive in shared memory by TMA bulk copies (cp.async.bulk, one elected thread,
//             completion on an mbarrier) -- issued for batch i + 1 BEFORE batch i is worked on: two
//             stages, the copy of the next batch always in flight.  The batch descriptors themselves
//             (32 B) travel one more iteration ahead in registers.
//   scatter   flat over the staged records, four per thread (two LDS.128); a record's read comes
//             from the staged CSR (a per-batch table gives the read of every 32nd record), the code
//             is the same for every lane also where a group of four crosses a read boundary.  Events
//             as in the first form: four packed +-1 per record.
//   scan      as in the first form: one block-wide prefix sum over the batch's packed histogram;
//             the scanned words leave as 256-bit stores (one per thread, fully coalesced).
//
// Bytes per record from HBM: 8 (abpos, aepos) instead of 12.  Measured: correct, long-scoreboard stalls
// gone, and still SLOWER than the other two forms (0.39 vs 0.32 ms on the short-read set, 1.05 vs 0.78 ms
// on the long-read set): two 512-thread CTAs per SM (107 KB of staging each) whose phases cannot overlap
// the way five to seven small CTAs do, and the four-event scatter on the same shared-memory data pipe
// (DESIGN.md section 4a).  Kept as a tested variant (HG_OPT_PROFILE_KERNEL = 3), not the default.
constexpr int kTmaThreads = 512;
constexpr int kTmaItems = kFlatBins / kTmaThreads;  // 8 bins per thread in the scan
constexpr int kTmaTab = kFlatMaxReads + 8;          // table entries per stage (+ alignment slack)

struct __align__(128) TmaStage {
    int s[kSlabRecords + 8];        // abpos of the batch's records, from the 16-byte boundary below the first
    int e[kSlabRecords + 8];        // aepos
    long long off[kTmaTab];         // read_off[f0 & ~1 ...]
    int base[kTmaTab];              // rbase[f0 & ~3 ...]
    int rlen[kTmaTab];
};
struct __align__(128) TmaShared {
    TmaStage st[2];
    uint32_t hist[kFlatWords];
    int sum[kFlatMaxReads], mx[kFlatMaxReads], mx2[kFlatMaxReads];
    int o32[kFlatMaxReads + 8];     // first staged record of every read of the batch
    int blk[(kSlabRecords + 8) / 32 + 4];  // read of the first record of every block of 32 staged records
    uint32_t wtot[kTmaThreads / 32];
    unsigned long long full[2];     // mbarriers: stage filled
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// global -> shared bulk copy (TMA, 1-D): 16-byte aligned addresses, size a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// The histogram of this form is not swizzled: the two LDS.128 per thread of the scan take a 2-way
// bank conflict (32 extra wavefronts per batch), cheaper than two more instructions in front of
// each of the ~15 000 atomics of a batch.
// bin of an event at position x >= -40 on the 40-bp grid: cov_bin(x) for x >= 0, 0 below
__device__ __forceinline__ unsigned tma_bin(int x) { return (unsigned)(x + kReso) / (unsigned)kReso; }

struct TmaBatch {
    int f0, nreads, words;          // desc[2 b]
    int64_t k0;                     // first record
    int nrec;
};
__device__ __forceinline__ TmaBatch tma_batch(const FlatParams& F, int b) {
    TmaBatch t;
    if (b < F.nbatch) {
        const int4 x = __ldg(F.desc + 2 * (size_t)b), y = __ldg(F.desc + 2 * (size_t)b + 1);
        t.f0 = x.x; t.nreads = x.y; t.words = x.z;
        t.k0 = (int64_t)(((unsigned long long)(unsigned)y.y << 32) | (unsigned)y.x);
        t.nrec = y.z;
    } else {
        t.f0 = 0; t.nreads = 0; t.words = 0; t.k0 = 0; t.nrec = 0;
    }
    return t;
}

// Issued by ONE thread: everything batch `t` needs, into stage `sg`, completion on `bar`.
__device__ __forceinline__ void tma_issue(const RecView& rv, const ReadView& rd, const FlatParams& F,
                                          const TmaBatch& t, TmaStage* sg, unsigned long long* bar) {
    // order the generic-proxy reads of this stage (two iterations ago, behind a CTA barrier) before
    // the async-proxy writes
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int fa = t.f0 & ~1, fb = t.f0 & ~3;
    const uint32_t n_off = (uint32_t)((t.f0 + t.nreads + 1 - fa + 1) & ~1);
    const uint32_t n_tab = (uint32_t)((t.f0 + t.nreads - fb + 3) & ~3);
    int64_t ka = 0;
    uint32_t n4 = 0;
    int tail = 0;
    if (t.words > 0 && t.nrec > 0) {
        ka = t.k0 & ~(int64_t)3;
        const int span = (int)(t.k0 - ka) + t.nrec;   // records from the aligned start
        n4 = (uint32_t)((span + 3) & ~3);
        if (ka + (int64_t)n4 > (rv.novl & ~(int64_t)3)) {  // the array's last records: no 16-byte multiple left
            n4 = (uint32_t)(span & ~3);
            tail = span - (int)n4;
        }
    }
    const uint32_t bytes = 8u * n4 + 8u * n_off + 8u * n_tab;
    mbar_expect_tx(bar, bytes);
    if (n4) {
        bulk_g2s(sg->s, rv.abpos + ka, 4u * n4, bar);
        bulk_g2s(sg->e, rv.aepos + ka, 4u * n4, bar);
    }
    bulk_g2s(sg->off, rv.read_off + fa, 8u * n_off, bar);
    if (n_tab) {
        bulk_g2s(sg->base, F.rbase + fb, 4u * n_tab, bar);
        bulk_g2s(sg->rlen, rd.rlen + fb, 4u * n_tab, bar);
    }
    for (int i = 0; i < tail; i++) {  // at most once per launch
        sg->s[n4 + i] = rv.abpos[ka + n4 + i];
        sg->e[n4 + i] = rv.aepos[ka + n4 + i];
    }
}

__global__ void __launch_bounds__(kTmaThreads, 2)
k_profile_tma(RecView rv, ReadView rd, hg_filter_params P, FlatParams F) {
    extern __shared__ __align__(128) unsigned char tma_smem_raw[];
    TmaShared& sh = *reinterpret_cast<TmaShared*>(tma_smem_raw);
    constexpr int NW = kTmaThreads / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x;
    const int C = P.cut_off;

    if (tid == 0) {
        mbar_init(&sh.full[0], 1);
        mbar_init(&sh.full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    TmaBatch cur = tma_batch(F, blockIdx.x), nxt = tma_batch(F, blockIdx.x + G);
    if (tid == 0 && blockIdx.x < F.nbatch) tma_issue(rv, rd, F, cur, &sh.st[0], &sh.full[0]);

    int it = 0;
    for (int b = blockIdx.x; b < F.nbatch; b += G, it++) {
        TmaStage& sg = sh.st[it & 1];
        // ---- the next batch's copies go out first; its descriptor came an iteration ago, the one
        // after it is requested now
        if (tid == 0 && b + G < F.nbatch) tma_issue(rv, rd, F, nxt, &sh.st[(it + 1) & 1], &sh.full[(it + 1) & 1]);
        const TmaBatch nn = tma_batch(F, b + 2 * G);
        const int nb = cur.words, nreads = cur.nreads;
        // per-read inputs of the last phase, requested now
        int my_self = 0;
        if (tid < nreads) my_self = __ldg(F.self_cnt + cur.f0 + tid);

        // ---- zero (the scan of the previous batch has read the histogram: its barrier lies between)
        if (nb > 0) {
            const int nz = min(kFlatWords, (nb + 16 + 3) & ~3);
            for (int j = tid * 4; j < nz; j += kTmaThreads * 4) sts128(sh.hist + j, make_uint4(0, 0, 0, 0));
        }
        if (tid < nreads) {   // kFlatMaxReads <= kTmaThreads: read tid of the batch belongs to thread tid throughout
            sh.sum[tid] = 0;
            sh.mx[tid] = -1;
        }

        // ---- wait for this batch's stage
        {
            const uint32_t parity = (uint32_t)(it >> 1) & 1u;
            if (!mbar_try_wait(&sh.full[it & 1], parity)) {
                const unsigned long long t0 = global_timer_ns();
                while (!mbar_try_wait(&sh.full[it & 1], parity))
                    if (global_timer_ns() - t0 > 2000000000ull) __trap();  // a lost copy must not hang the GPU
            }
        }
        const int d_off = cur.f0 & 1, d_tab = cur.f0 & 3;   // where read f0 sits in the tables
        const int64_t ka = cur.k0 & ~(int64_t)3;

        // ---- per batch: record offsets of the reads inside the stage, and for every block of 32
        // staged records the read its first record belongs to (one writer per block); the read's
        // table entries move to registers (the stage is refilled before the last phase)
        const int rec_lo = (int)(cur.k0 - ka), rec_hi = rec_lo + cur.nrec;
        int my_base = -1, my_rl = 0;
        if (tid == 0) sh.o32[nreads] = rec_hi;   // = read_off[f0 + nreads] - ka: every read of a batch with words fits it
        if (tid < nreads) {
            const int lo = (int)(sg.off[d_off + tid] - ka);
            sh.o32[tid] = lo;
            my_base = sg.base[d_tab + tid];
            my_rl = sg.rlen[d_tab + tid];
            const int hi = (int)(sg.off[d_off + tid + 1] - ka);
            if (hi > lo && nb > 0) {
                if (lo == rec_lo) sh.blk[lo >> 5] = tid;
                for (int m = (lo + 31) >> 5; (m << 5) < hi; m++) sh.blk[m] = tid;
            }
        }
        __syncthreads();

        // ---- scatter (profileCoverage, LAInterface.cpp:4298-4320): flat over the staged records,
        // four per thread and step (two LDS.128).  Every record is scattered, A == B ones included
        // (telling them apart would need the bread column): they are taken out again below.
        if (nb > 0 && cur.nrec > 0) {
            // The lanes of a warp are spread over eight windows (see flat_group): many records start
            // in their read's first bin or end in its last one, and the four lanes of a window are all
            // that can meet on such a word.  A window is WS groups of four records, the batch's groups
            // divided by eight (a multiple of 8 in 32 .. 64); the position inside the window is rotated
            // by four groups per window, which keeps the 128-bit loads of a quarter warp on eight
            // different bank groups.
            const int NG = (rec_hi + 3) >> 2;
            const int WS = NG >= 512 ? 64 : max(32, (((NG + 7) >> 3) + 7) & ~7);
            const int wdw = lane >> 2;
            int pos = 4 * warp + (lane & 3);
            const bool busy = pos < WS;
            pos += 4 * wdw;
            if (pos >= WS) pos -= WS;
            for (int g = 4 * (wdw * WS + pos); busy && g < rec_hi; g += 32 * WS) {
                const int4 vs = *reinterpret_cast<const int4*>(sg.s + g), ve = *reinterpret_cast<const int4*>(sg.e + g);
                int r0 = sh.blk[g >> 5];
                int hi0 = sh.o32[r0 + 1];
                const int kf = max(g, rec_lo);
                while (kf >= hi0) {   // the read of the group's first record
                    r0++;
                    hi0 = sh.o32[r0 + 1];
                }
                uint32_t* hb0 = sh.hist + sg.base[d_tab + r0];
                uint32_t* hb1 = hb0;
                const int nA = hi0 - g;   // records of the group that belong to read r0 (when all four are in the batch)
                int r1 = r0;
                bool simple = g >= rec_lo && g + 4 <= rec_hi;
                if (simple && nA < 4) {   // a read boundary inside the group: the next read that has records
                    int hi1;
                    do {
                        r1++;
                        hi1 = sh.o32[r1 + 1];
                    } while (hi1 <= hi0);
                    hb1 = sh.hist + sg.base[d_tab + r1];
                    simple = hi1 >= g + 4;
                }
                if (simple) {
                    // the same code for every lane: records i < nA go to read r0, the others to r1
                    const int sv[4] = {vs.x, vs.y, vs.z, vs.w}, ev[4] = {ve.x, ve.y, ve.z, ve.w};
                    int accT = 0, accA = 0, mxA = 0, mxB = 0;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const bool inA = i < nA;
                        uint32_t* const hb = inA ? hb0 : hb1;
                        const unsigned q_s = tma_bin(sv[i]), q_e = tma_bin(ev[i]);   // 0 <= abpos < aepos <= rlen (ingest check)
                        const unsigned b_sc = tma_bin(max(sv[i] + C, -kReso)), b_ec = tma_bin(max(ev[i] - C, -kReso));
                        atomicAdd(hb + q_s, 1u);
                        atomicAdd(hb + q_e, 0u - 1u);
                        atomicAdd(hb + b_sc, 1u << 16);
                        atomicAdd(hb + b_ec, 0u - (1u << 16));
                        const int d = (int)q_e - (int)q_s;
                        accT += d;
                        accA += inA ? d : 0;
                        mxA = max(mxA, inA ? (int)q_e : 0);
                        mxB = max(mxB, inA ? 0 : (int)q_e);
                    }
                    atomicAdd(&sh.sum[r0], accA);
                    atomicMax(&sh.mx[r0], mxA);
                    if (nA < 4) {
                        atomicAdd(&sh.sum[r1], accT - accA);
                        atomicMax(&sh.mx[r1], mxB);
                    }
                    continue;
                }
                // a group at an end of the batch, or one that touches more than two reads
                const int sv[4] = {vs.x, vs.y, vs.z, vs.w}, ev[4] = {ve.x, ve.y, ve.z, ve.w};
                int r = r0, hi = hi0;
                uint32_t* hb = hb0;
                int acc = 0, mxq = -1;
                for (int i = 0; i < 4; i++) {
                    const int k = g + i;
                    if (k < rec_lo || k >= rec_hi) continue;
                    if (k >= hi) {   // next read (reads without records are skipped)
                        atomicAdd(&sh.sum[r], acc);
                        atomicMax(&sh.mx[r], mxq);
                        acc = 0;
                        mxq = -1;
                        do {
                            r++;
                            hi = sh.o32[r + 1];
                        } while (k >= hi);
                        hb = sh.hist + sg.base[d_tab + r];
                    }
                    const unsigned q_s = tma_bin(sv[i]), q_e = tma_bin(ev[i]);
                    const unsigned b_sc = tma_bin(max(sv[i] + C, -kReso)), b_ec = tma_bin(max(ev[i] - C, -kReso));
                    atomicAdd(hb + q_s, 1u);
                    atomicAdd(hb + q_e, 0u - 1u);
                    atomicAdd(hb + b_sc, 1u << 16);
                    atomicAdd(hb + b_ec, 0u - (1u << 16));
                    acc += (int)q_e - (int)q_s;
                    mxq = max(mxq, (int)q_e);
                }
                if (mxq >= 0) {
                    atomicAdd(&sh.sum[r], acc);
                    atomicMax(&sh.mx[r], mxq);
                }
            }
            // A == B records are inactive (filter.cpp:538-547) and rare; their per-read count comes
            // from the ingest.  One thread per such read takes their events out again (atomic adds
            // commute) and finds the maximum without them.
            if (tid < nreads && self_count(my_self) > 0) {
                const int read = cur.f0 + tid;
                const int base = my_base;
                if (base >= 0) {
                    int mx = -1, acc = 0;
                    for (int64_t k = rv.read_off[read]; k < rv.read_off[read + 1]; k++) {
                        const int as = rv.abpos[k], ae = rv.aepos[k];
                        const int q_s = as / kReso + 1, q_e = ae / kReso + 1;
                        if (rv.bread[k] != read) {
                            mx = max(mx, q_e);
                            continue;
                        }
                        atomicAdd(&sh.hist[base + q_s], 0u - 1u);
                        atomicAdd(&sh.hist[base + q_e], 1u);
                        atomicAdd(&sh.hist[base + cov_bin(as + C, kReso)], 0u - (1u << 16));
                        atomicAdd(&sh.hist[base + cov_bin(ae - C, kReso)], 1u << 16);
                        acc += q_e - q_s;
                    }
                    atomicAdd(&sh.sum[tid], -acc);
                    sh.mx2[tid] = mx;
                }
            }
        }
        __syncthreads();

        // ---- one prefix sum over the whole batch; the scanned words go to HBM for K2
        if (nb > 0) {
            uint32_t* const pw = F.prof + (size_t)b * kFlatBins;
            const int j0 = tid * kTmaItems;
            uint32_t v[kTmaItems];
            if (j0 < nb) {
#pragma unroll
                for (int q = 0; q < kTmaItems / 4; q++) {
                    const uint4 x = lds128(sh.hist + j0 + 4 * q);
                    v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
                }
#pragma unroll
                for (int i = 1; i < kTmaItems; i++) v[i] += v[i - 1];
            } else {
#pragma unroll
                for (int i = 0; i < kTmaItems; i++) v[i] = 0;
            }
            const uint32_t incl = warp_incl_scan(v[kTmaItems - 1]);
            if (lane == 31) sh.wtot[warp] = incl;
            __syncthreads();
            // the 16 warp totals: one more warp-level scan (every warp does its own, no second barrier)
            uint32_t wt = lane < NW ? sh.wtot[lane] : 0u;
            wt = warp_incl_scan(wt);
            const uint32_t wpre = __shfl_sync(0xffffffffu, wt, (warp + 31) & 31);  // inclusive total of warp - 1
            const uint32_t pre = incl - v[kTmaItems - 1] + (warp > 0 ? wpre : 0u);
            if (j0 < nb) {
#pragma unroll
                for (int i = 0; i < kTmaItems; i++) v[i] += pre;
                st_global_v8(pw + j0, v);
            }
        }

        // ---- per read: length and mean of the cut-off-free profile (filter.cpp:642-656).  No barrier
        // closes the iteration: what the next one overwrites early (histogram, the other stage's
        // tables) was last read before the scan's barrier, and a read's sum / maximum are reset by
        // the thread that reads them here.
        if (tid < nreads) {
            const int read = cur.f0 + tid;
            if (my_base < 0 || nb == 0)
                F.big_list[atomicAdd(&F.counters1[0], 1)] = read;
            else
                finalize_read(rv, F, read, sh.sum[tid], self_count(my_self) > 0 ? sh.mx2[tid] : sh.mx[tid], my_rl, my_self);
        }
        cur = nxt;
        nxt = nn;
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
//
// Two kernels.  k_mask_bits_flat streams the stored profiles batch by batch and leaves two bits per
// bin in HBM; k_mask_walk then gives every read ONE THREAD, 32 reads to a warp.  A read's walk is a
// chain of dependent steps (bit-map words, coverage look-ups behind the annotation candidates); in a
// single kernel it ran on the first lanes of one warp of the batch's CTA while the CTA's other
// warps had retired but still held their slots: eight walking warps per SM, 0.61 ms on the long-read
// set.  Measured alternatives in that form: reads dealt out over the CTA's warps 1.16 ms (the walk
// is ~1500 instructions, so six instruction streams per CTA cost six times the issue slots), a
// warp-wide walk (lanes over bit-map words) 0.95 ms.

// Bit i of the word <=> entry 32 w + i; restricted to entries in [lo, hi).
__device__ __forceinline__ uint32_t map_clip(uint32_t m, int w, int lo, int hi) {
    const int b = w << 5;
    if (lo > b) m &= lo - b >= 32 ? 0u : (0xffffffffu << (lo - b));
    if (hi < b + 32) m &= hi <= b ? 0u : (0xffffffffu >> (b + 32 - hi));
    return m;
}

template <bool DUMP>
__global__ void __launch_bounds__(kFlatThreads, 8)
k_mask_bits_flat(RecView rv, hg_filter_params P, FlatParams F, MaskAnnoOut out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int2 bt = F.batch[blockIdx.x];
    const int f0 = bt.x, f1 = F.batch[blockIdx.x + 1].x;
    const int MIN_COV = F.scal[1];
    // MIN_COV < 0: runs could cross read boundaries, everything goes the generic way; so do the
    // batches the second form of K1 declined (more records than its 16-bit prefix counts hold)
    const bool declined = F.v2 == 1 && bt.y > 0 && rv.read_off[f1] - rv.read_off[f0] > kV2MaxBatchRecords;
    const int nb = (MIN_COV < 0 || declined) ? 0 : bt.y;
    const int npass = (nb + kFlatPass - 1) / kFlatPass;
    const uint32_t* __restrict__ const pw = F.prof + (size_t)blockIdx.x * kFlatBins;
    uint16_t* const zmap16 = F.zmap + (size_t)blockIdx.x * kFlatMaps;
    uint16_t* const cmap16 = F.cmap + (size_t)blockIdx.x * kFlatMaps;

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
            for (int q = 0; q < kFlatItems / 8; q++) ld_global_nc_v8(pw + j0 + 8 * q, v + 8 * q);
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

constexpr int kWalkThreads = 128;
constexpr int kWalkFetch = 4;   // bit-map words in flight per thread

// Per-read state of the walk between its phases.
struct WalkRead {
    const uint32_t* pw;     // the batch's packed profiles
    const uint32_t* cmap;
    int base, L0, ja_lo, ja_hi, MIN_COV;
};

// Repeat annotation from the coverage gradient (filter.cpp:796-813) + merge pass (filter.cpp:817-829)
// as a stream over the flagged bins; WRITE = false counts what survives, WRITE = true stores it.
// Candidates are bins j < L0 - 2 with 40 j in [mask.start + NHR, mask.end - NHR]; map entry j + 1
// flags the jump cov0[j + 1] - cov0[j].
template <bool WRITE>
__device__ __forceinline__ int walk_annotations(const WalkRead& w, const hg_filter_params& P, const MaskAnnoOut& out,
                                                int off) {
    constexpr int reso = kReso;
    const int MINT = P.min_repeat_annotation_threshold, MAXT = P.max_repeat_annotation_threshold;
    const int GAP = P.repeat_annotation_gap_threshold;
    int n = 0;
    unsigned cur = 0;
    bool have = false;
    if (w.ja_hi >= w.ja_lo) {
        const int lo = w.base + w.ja_lo + 1, hi = w.base + w.ja_hi + 2;
        const int w1 = (hi - 1) >> 5;
        const int wf = lo >> 5;
        const uint32_t mf = 0xffffffffu << (lo & 31), ml = 0xffffffffu >> (31 - ((hi - 1) & 31));  // clips of the first / last word
        for (int w0 = wf; w0 <= w1; w0 += kWalkFetch) {
            uint32_t buf[kWalkFetch], any = 0;
#pragma unroll
            for (int i = 0; i < kWalkFetch; i++) {
                buf[i] = w0 + i <= w1 ? __ldg(w.cmap + w0 + i) : 0u;
                any |= buf[i];
            }
            if (any == 0) continue;   // most words flag nothing
#pragma unroll
            for (int i = 0; i < kWalkFetch; i++) {
                const int wd = w0 + i;
                uint32_t m = buf[i];
                if (wd == wf) m &= mf;
                if (wd == w1) m &= ml;
                while (m) {
                    const int bit = __ffs(m) - 1;
                    m &= m - 1;
                    const int j = (wd << 5) + bit - 1 - w.base;
                    const int c0 = f_lo(__ldg(w.pw + w.base + j));
                    const int g = f_lo(__ldg(w.pw + w.base + j + 1)) - c0;
                    const int thr = min(max((c0 + w.MIN_COV) / P.coverage_fraction, MINT), MAXT);
                    const int type = g > thr ? 1 : (g < -thr ? -1 : 0);
                    if (type == 0) continue;
                    const unsigned nx = ((unsigned)(reso * j) << 2) | (unsigned)(type + 1);
                    if (!have) {
                        cur = nx;
                        have = true;
                        continue;
                    }
                    const int ct = (int)(cur & 3u) - 1;
                    const int gap = (int)(nx >> 2) - (int)(cur >> 2);
                    if (ct == 1 && type == 1 && gap < GAP) {
                        continue;   // +1,+1 close together: the later one goes
                    } else if (ct == -1 && type == -1 && gap < GAP) {
                        cur = nx;   // -1,-1 close together: the earlier one goes
                    } else {
                        if (WRITE) {
                            out.anno_pool[off + n] = make_int2((int)(cur >> 2), (int)(cur & 3u) - 1);
                            out.hinge_keep[off + n] = 0;
                        }
                        n++;
                        cur = nx;
                    }
                }
            }
        }
    }
    if (have) {
        if (WRITE) {
            out.anno_pool[off + n] = make_int2((int)(cur >> 2), (int)(cur & 3u) - 1);
            out.hinge_keep[off + n] = 0;
        }
        n++;
    }
    return n;
}

// One thread per read, 32 reads to a warp.  Pool space for the annotations and slots of the hinge work
// list are claimed ONCE PER WARP (prefix sums over the lanes' demands, one atomic by lane 0): a quarter
// of a million single-address atomics with return values would otherwise take as long as the walk.
__global__ void __launch_bounds__(kWalkThreads)
k_mask_walk(RecView rv, ReadView rd, hg_filter_params P, FlatParams F, MaskAnnoOut out) {
    const int lane = threadIdx.x & 31;
    const int read = F.p_lo + blockIdx.x * kWalkThreads + threadIdx.x;
    const int MIN_COV = F.scal[1];
    constexpr int reso = kReso;
    const int NHR = P.no_hinge_region;
    bool act = read < F.p_hi;
    const int bi = act ? F.rbatch[read] : -1;
    act = act && bi >= 0;
    int64_t nrec = 0;
    WalkRead w;
    w.pw = nullptr; w.cmap = nullptr; w.base = -1; w.L0 = 0; w.ja_lo = 0; w.ja_hi = -1; w.MIN_COV = MIN_COV;
    if (act) {
        w.base = F.rbase[read];
        nrec = rv.read_off[read + 1] - rv.read_off[read];
        bool generic = w.base < 0 || MIN_COV < 0 || nrec > Packed<uint32_t>::kMaxCount;
        if (!generic && F.v2 == 1) {  // a batch the second form of K1 declined
            const int f0 = F.batch[bi].x, f1 = F.batch[bi + 1].x;
            generic = rv.read_off[f1] - rv.read_off[f0] > kV2MaxBatchRecords;
        }
        if (generic) {
            out.big_list[atomicAdd(&out.counters[3], 1)] = read;
            act = false;
        }
    }

    // ---- phase 1: mask, telomere flag, number of annotations
    int2 mk = make_int2(0, 0);
    int msc = 0, mec = 0, kept = 0;
    uint8_t flags = 0;
    if (act) {
        const int base = w.base;
        w.pw = F.prof + (size_t)bi * kFlatBins;
        w.cmap = reinterpret_cast<const uint32_t*>(F.cmap + (size_t)bi * kFlatMaps);
        const uint32_t* __restrict__ const zmap = reinterpret_cast<const uint32_t*>(F.zmap + (size_t)bi * kFlatMaps);
        const int nbz = bins_needed(rd.rlen[read], P);
        w.L0 = F.cov_maxbin[read] + 1;  // length of the cut-off-free profile
        auto H = [&](int j) { return __ldg(w.pw + base + j); };  // packed coverage of the read's bin j

        // longest run of covered bins (filter.cpp:696-728): the run between two consecutive zeros
        // p < z scores 40 (z - p - 2); bin 0 acts as a zero; '>' keeps the earliest of the longest
        int p = base, bestgap = 0, bestz = 0;
        const int end = base + nbz;
        {
            // the map words of a read are fetched kWalkFetch at a time (independent loads)
            const int w1 = (end - 1) >> 5, wf = (base + 1) >> 5;
            const uint32_t mf = 0xffffffffu << ((base + 1) & 31), ml = 0xffffffffu >> (31 - ((end - 1) & 31));
            for (int w0 = wf; w0 <= w1; w0 += kWalkFetch) {
                uint32_t buf[kWalkFetch], any = 0;
#pragma unroll
                for (int i = 0; i < kWalkFetch; i++) {
                    buf[i] = w0 + i <= w1 ? __ldg(zmap + w0 + i) : 0u;
                    any |= buf[i];
                }
                if (any == 0) continue;   // a covered stretch: no zeros in these words
#pragma unroll
                for (int i = 0; i < kWalkFetch; i++) {
                    const int wd = w0 + i;
                    uint32_t m = buf[i];
                    if (wd == wf) m &= mf;
                    if (wd == w1) m &= ml;
                    while (m) {
                        const int bit = __ffs(m) - 1;
                        const int z = (wd << 5) + bit;
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
            }
        }
        int maxstart = 0, maxend = 0;
        if (bestgap >= 3) {
            const int z = bestz - base, pz = z - bestgap;
            msc = pz + 1;
            mec = z - 1;
            maxstart = reso * (pz + 1);
            maxend = reso * (z - 1);
        }

        // telomere / coverage-imbalance flag (filter.cpp:731-760)
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
        if (P.use_qv_mask && P.use_coverage_mask)
            mk = make_int2(max(maxstart, q.x), min(maxend, q.y));
        else if (P.use_coverage_mask && !P.use_qv_mask)
            mk = make_int2(maxstart, maxend);
        else
            mk = q;

        w.ja_lo = mk.x + NHR <= 0 ? 0 : (mk.x + NHR + reso - 1) / reso;
        w.ja_hi = mk.y - NHR < 0 ? -1 : min((mk.y - NHR) / reso, w.L0 - 3);
        kept = walk_annotations<false>(w, P, out, 0);
    }

    // ---- pool space for the warp's annotations: one atomic
    int off = 0;
    {
        int incl = kept;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        int first = 0;
        if (total > 0) {
            if (lane == 0) first = atomicAdd(&out.counters[0], total);
            first = __shfl_sync(0xffffffffu, first, 0);
            if (first + total > out.anno_cap) {   // the host grows the pool and reruns the stage
                if (lane == 0) atomicExch(&out.counters[2], 1);
                first = -1;
            }
        }
        off = first < 0 ? -1 : first + incl - kept;
    }

    // ---- phase 2: the annotations, the hinge pre-test, the read's results
    bool push = false;
    if (act) {
        if (kept > 0 && off >= 0) walk_annotations<true>(w, P, out, off);
        // hinge pre-test: mean coverage near both mask ends (filter.cpp:842-865); its outcome
        // only matters for reads that carry annotations
        bool skip_hinges = false;
        if (kept > 0) {
            auto H = [&](int j) { return __ldg(w.pw + w.base + j); };
            int cs = 0, ns = 0, ce = 0, ne = 0;
            int jlo = mk.x <= 0 ? 0 : (mk.x + reso - 1) / reso;  // bins with mk.x <= 40 j <= mk.x + NHR
            int jhi = mk.x + NHR < 0 ? -1 : min((mk.x + NHR) / reso, w.L0 - 1);
            for (int j = jlo; j <= jhi; j++) {
                cs += f_lo(H(j));
                ns++;
            }
            jlo = mk.y - NHR <= 0 ? 0 : (mk.y - NHR + reso - 1) / reso;  // mk.y - NHR <= 40 j <= mk.y
            jhi = mk.y < 0 ? -1 : min(mk.y / reso, w.L0 - 1);
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
        out.anno_ref[read] = make_int2(kept > 0 ? off : 0, kept);
        push = kept > 0 && !skip_hinges && off >= 0;
    }

    // ---- slots of the hinge work list: one atomic per warp
    {
        const unsigned m = __ballot_sync(0xffffffffu, push);
        if (m) {
            int first = 0;
            if (lane == 0) first = atomicAdd(&out.counters[1], __popc(m));
            first = __shfl_sync(0xffffffffu, first, 0);
            if (push) write_work_item(out, first + __popc(m & ((1u << lane) - 1u)), read, rv.read_off[read], (int)nrec, mk, off, kept);
        }
    }
}

// ------------------------------------------------------------------ host side

// Greedy packing of the reads [lo, hi) into batches of at most kFlatBins histogram words,
// kFlatMaxReads reads and -- with cap_records -- kSlabRecords records (what the third form of K1
// stages in shared memory per batch).  A read that is too long or too deep for a batch goes to
// the per-read fallbacks (rbase = -1) and is a batch of its own (words = 0), which reports it.
void flat_plan(const int* rlen, const int64_t* read_off, int lo, int hi, int n_read, int cut_off, bool cap_records,
               FlatPlan* plan) {
    const int64_t max_recs = cap_records && read_off ? kSlabRecords : (int64_t)1 << 60;
    plan->batch.clear();
    plan->desc.clear();
    plan->rbase.assign((size_t)n_read, -1);
    plan->rbatch.assign((size_t)n_read, -1);
    int used = 0;
    int64_t recs = 0;
    bool open = false, solitary = false;
    auto close = [&](int r_end) {
        if (!open) return;
        const int f0 = plan->batch.back().x;
        const int64_t k0 = read_off ? read_off[f0] : 0;
        plan->desc.push_back(make_int4(f0, r_end - f0, plan->batch.back().y, 0));
        plan->desc.push_back(make_int4((int)(k0 & 0xffffffffll), (int)(k0 >> 32), (int)recs, 0));
        open = false;
    };
    for (int r = lo; r < hi; r++) {
        const int nbz = bins_needed(rlen[r], cut_off);
        const int64_t nrec = read_off ? read_off[r + 1] - read_off[r] : 0;
        const bool fits = nbz <= kFlatBins && nrec <= max_recs;
        if (!open || solitary || !fits || used + nbz > kFlatBins || recs + nrec > max_recs ||
            r - plan->batch.back().x >= kFlatMaxReads) {
            close(r);
            plan->batch.push_back(make_int2(r, 0));
            open = true;
            solitary = !fits;
            used = 0;
            recs = 0;
        }
        plan->rbatch[r] = (int)plan->batch.size() - 1;
        if (!fits) continue;
        plan->rbase[r] = used;
        used += nbz;
        recs += nrec;
        plan->batch.back().y = used;
    }
    if (!open) {
        plan->batch.push_back(make_int2(lo, 0));
        open = true;
    }
    close(hi);
    plan->batch.push_back(make_int2(hi, 0));
}

// The second form of K1 is written for the nominal cut-off (300: every INI the reference ships).
// It halves the shared-memory atomics per record but handles two histogram words per coverage bin,
// so it only pays where records outnumber bins (short reads, deep pile-ups: 0.354 vs 0.347 ms on
// the 62 M-record N(3500,1500) set, a wash); with long reads (3 bins per record on the
// N(24000,8000) set) the first form is faster (0.92 vs 1.14 ms).  flat_bins_per_record is set
// when the batch plan is made.
// The third form (k_profile_tma) needs 16-byte aligned columns for its bulk copies.
static bool use_v3(const FilterScratch& s, const RecView& rv) {
    if (s.flat_kernel != 3 || !s.flat_capped) return false;
    return ((reinterpret_cast<uintptr_t>(rv.abpos) | reinterpret_cast<uintptr_t>(rv.aepos)) & 15) == 0;
}

static bool use_v2(const FilterScratch& s, const hg_filter_params& P) {
    if (s.flat_kernel == 1 || s.flat_kernel == 3 || P.cut_off != kV2CutOff) return false;
    if (s.flat_kernel >= 5) return true;
    return s.flat_bins_per_record < 1.5f;
}

static FlatParams flat_params(const FilterScratch& s, const hg_filter_params& P, const RecView& rv, int r_begin,
                              int r_end) {
    FlatParams F;
    F.v2 = use_v3(s, rv) ? 2 : use_v2(s, P) ? 1 : 0;
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
    F.rbatch = s.flat_rbatch;
    F.desc = s.flat_desc;
    F.zmap = s.flat_zmap;
    F.cmap = s.flat_cmap;
    F.p_lo = s.flat_lo;
    F.p_hi = s.flat_hi;
    F.nbatch = s.flat_nbatch;
    F.r_begin = r_begin;
    F.r_end = r_end;
    return F;
}

// Phase 1 of the stage: both coverage profiles of every owned read, their lengths and means.
void launch_profile(const RecView& rv, const ReadView& rd, const hg_filter_params& P, int r_begin,
                    int r_end, FilterScratch& s, cudaStream_t st) {
    const FlatParams F = flat_params(s, P, rv, r_begin, r_end);
    // the counters of both phases; per-read results of reads outside the planned range were
    // cleared when the plan was made (hg_capi.cu)
    cudaMemsetAsync(s.counters, 0, sizeof(int) * 16, st);
    const int grid = s.flat_nbatch;
    if (grid <= 0) return;
    g_launches += 2;
    if (F.v2 == 2) {
        // persistent: two CTAs per SM, batches strided over them; function attributes are per device
        const int smem = (int)sizeof(TmaShared);
        cudaFuncSetAttribute(k_profile_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        k_profile_tma<<<std::min(grid, 2 * s.num_sms), kTmaThreads, smem, st>>>(rv, rd, P, F);
    } else if (F.v2) {
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
    const FlatParams F = flat_params(s, P, rv, r_begin, r_end);
    const int grid = s.flat_nbatch;
    if (grid <= 0) return;
    g_launches += 2;
    if (out.cov0)
        k_mask_bits_flat<true><<<grid, kFlatThreads, 0, st>>>(rv, P, F, out);
    else
        k_mask_bits_flat<false><<<grid, kFlatThreads, 0, st>>>(rv, P, F, out);
    const int nreads = s.flat_hi - s.flat_lo;
    if (nreads > 0)
        k_mask_walk<<<(nreads + kWalkThreads - 1) / kWalkThreads, kWalkThreads, 0, st>>>(rv, rd, P, F, out);
}

}  // namespace hg
