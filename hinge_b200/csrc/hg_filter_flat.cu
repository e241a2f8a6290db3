// K2, flat form: coverage mask + repeat annotation + hinge pre-test of the
// `hinge filter` stage (/root/reference/src/filter/filter.cpp:696-865,
// /root/reference/src/lib/LAInterface.cpp:4298-4320) for a BATCH of consecutive
// A-reads per CTA instead of one read per warp.
//
// Why this shape.  A PacBio-like read has ~90 coverage bins and ~90 pile-up records:
// far too little work for a warp, so a warp-per-read kernel spends its time in per-read
// fixed overhead executed by 32 mostly idle lanes (ncu: 606 warp instructions per read,
// 65 % issue-active, 11 % of HBM).  Here the profiles of ~40 reads are laid end to end in
// one shared-memory array and every phase runs flat over it with all lanes busy:
//
//   scatter   flat over the batch's records (int4 loads of aread/abpos/aepos, the next
//             step's loads in flight while this step's events go out): four packed +-1
//             events per record (low half: profile without cut-off, high half: with).
//             The lanes of a warp are spread over eight record windows (flat_group) so
//             that the many records that start in their read's first bin or end in its
//             last one do not all serialise on one shared-memory word.
//   scan      ONE block-wide prefix sum over the concatenated array.  Every record adds
//             +1 and -1 inside its own read's bins, so the running sum is back at zero at
//             every read boundary: no segmentation needed.  The same pass leaves two bit
//             maps: bins whose cut-off coverage is <= MIN_COV ("zeros") and bins where the
//             coverage jumps by more than the smallest annotation threshold.
//   per read  one thread per read walks its slice of the two bit maps: longest covered run
//             (filter.cpp:696-728), mask, telomere flag, repeat annotations with the
//             streaming form of the merge pass (filter.cpp:796-829), hinge pre-test
//             (filter.cpp:842-865).
//
// The kernel is bound by the shared-memory data pipe (ncu: l1tex data-pipe wavefronts > 80 %
// of peak): ~4 wavefronts per ATOMS on random bins is what the banks give, so the rest of the
// design keeps every other shared-memory access conflict-free (sw) and the global loads wide.
//
// Reads longer than kFlatBins bins, pile-ups deeper than the 16-bit halves can count and
// runs with MIN_COV < 0 go to the generic per-read kernel (k_mask_anno_big, hg_filter.cu).
#include "hg_device.cuh"
#include "hg_filter.h"

namespace hg {

extern int64_t g_launches;

constexpr int kFlatThreads = 256;
constexpr int kFlatItems = 16;                           // bins per thread and scan pass
constexpr int kFlatPass = kFlatThreads * kFlatItems;     // bins per scan pass
constexpr int kFlatWords = kFlatBins + 32;               // histogram words per CTA (+ slack for the j + 1 reads)
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

// The four events of one record (profileCoverage, LAInterface.cpp:4298-4320): low half of the
// packed word = profile without cut-off, high half = with.
__device__ __forceinline__ void scatter_record(uint32_t* hist, int base, int as, int ae, int C) {
    const int b_s0 = base + as / kReso + 1, b_e0 = base + ae / kReso + 1;  // 0 <= abpos < aepos (ingest check)
    const int b_sc = base + cov_bin(as + C, kReso), b_ec = base + cov_bin(ae - C, kReso);
    atomicAdd(&hist[sw(b_s0)], 1u);
    atomicAdd(&hist[sw(b_e0)], 0u - 1u);
    atomicAdd(&hist[sw(b_sc)], 1u << 16);
    atomicAdd(&hist[sw(b_ec)], 0u - (1u << 16));
}

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
    const int2* __restrict__ batch;       // nbatch + 1: (first read, histogram words in use) of every batch
    const int* __restrict__ rbase;        // per read: first word of its profile inside its batch, -1 = generic path
    const int* __restrict__ cov_maxbin;   // K1: last bin of the cut-off-free profile (-1 = empty pile-up)
    const int* __restrict__ batch_self;   // K1: the batch holds records with A == B
    const int* __restrict__ scal;         // [1] = MIN_COV
};

// Which 4-record group of a 1024-record tile a thread takes.  With the identity map the 32
// lanes of a warp hold 128 consecutive records, i.e. one or two reads, and the ~16 of them that
// start in the read's first bin (or end in its last) serialise on one shared-memory word.
// Spreading the lanes over SPREAD windows 128 records apart (32 / SPREAD lanes, a 16 * 32 / SPREAD
// byte run, per window) divides that multiplicity by SPREAD while every window still reads whole
// sectors.  Measured on B200 (K2 alone, ms): 1 -> 0.404, 4 -> 0.365, 8 -> 0.368, 16 -> 0.395, 32 -> 0.459.
template <int SPREAD>
__device__ __forceinline__ int flat_group(int tid) {
    if (SPREAD <= 1) return tid;
    constexpr int L = 32 / SPREAD;  // lanes per window
    const int lane = tid & 31, warp = tid >> 5;
    return (lane % L) + L * warp + (kFlatThreads / SPREAD) * (lane / L);
}

template <int SPREAD, bool DUMP>
__global__ void __launch_bounds__(kFlatThreads)
k_mask_anno_flat(RecView rv, ReadView rd, hg_filter_params P, FlatParams F, MaskAnnoOut out) {
    __shared__ __align__(16) uint32_t hist[kFlatWords];
    __shared__ __align__(16) uint16_t zmap16[kFlatMaps];  // bin has cut-off coverage <= MIN_COV
    __shared__ __align__(16) uint16_t cmap16[kFlatMaps];  // |cov0[j] - cov0[j-1]| above the smallest threshold
    __shared__ uint32_t wtot[2][kFlatThreads / 32];
    const uint32_t* zmap = reinterpret_cast<const uint32_t*>(zmap16);
    const uint32_t* cmap = reinterpret_cast<const uint32_t*>(cmap16);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int2 bt = F.batch[blockIdx.x];
    const int f0 = bt.x, f1 = F.batch[blockIdx.x + 1].x;
    const int MIN_COV = F.scal[1];
    constexpr int reso = kReso;
    // MIN_COV < 0: runs could cross read boundaries, everything goes the generic way
    const int nb = MIN_COV < 0 ? 0 : bt.y;
    const int npass = (nb + kFlatPass - 1) / kFlatPass;
    const bool self_records = F.batch_self[blockIdx.x] != 0;  // otherwise the bread column is not even loaded

    // ---- the batch's records: four per thread and step
    const int64_t k_begin = nb > 0 ? rv.read_off[f0] : 0, k_end = nb > 0 ? rv.read_off[f1] : 0;
    const int64_t g0 = k_begin & ~(int64_t)3;
    const int grp = flat_group<SPREAD>(tid) * 4;
    int4 va = make_int4(-1, -1, -1, -1), vs = make_int4(0, 0, 0, 0), ve = vs, vb = make_int4(-2, -2, -2, -2);
    auto load4 = [&](int64_t k) {
        if (k >= k_begin && k + 4 <= k_end) {
            va = __ldg(reinterpret_cast<const int4*>(rv.aread + k));
            vs = __ldg(reinterpret_cast<const int4*>(rv.abpos + k));
            ve = __ldg(reinterpret_cast<const int4*>(rv.aepos + k));
            if (self_records) vb = __ldg(reinterpret_cast<const int4*>(rv.bread + k));
        } else {  // ragged ends of the batch: records outside it get read id -1
            int a[4], s[4], e[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int64_t ki = k + i;
                const bool in = ki >= k_begin && ki < k_end;
                a[i] = in ? __ldg(rv.aread + ki) : -1;
                s[i] = in ? __ldg(rv.abpos + ki) : 0;
                e[i] = in ? __ldg(rv.aepos + ki) : 0;
                b[i] = (in && self_records) ? __ldg(rv.bread + ki) : -2;
            }
            va = make_int4(a[0], a[1], a[2], a[3]);
            vs = make_int4(s[0], s[1], s[2], s[3]);
            ve = make_int4(e[0], e[1], e[2], e[3]);
            vb = make_int4(b[0], b[1], b[2], b[3]);
        }
    };
    if (g0 < k_end) load4(g0 + grp);  // in flight while the histogram is cleared

    // ---- zero
    for (int j = tid * 4; j < npass * kFlatPass + 16 && j < kFlatWords; j += kFlatThreads * 4)
        sts128(hist + j, make_uint4(0, 0, 0, 0));
    __syncthreads();

    // ---- scatter: the next step's loads are issued before this step's events go out
    {
        const int C = P.cut_off;
        for (int64_t kb = g0; kb < k_end; kb += kFlatThreads * 4) {
            const int a[4] = {va.x, va.y, va.z, va.w}, s[4] = {vs.x, vs.y, vs.z, vs.w};
            const int e[4] = {ve.x, ve.y, ve.z, ve.w}, b[4] = {vb.x, vb.y, vb.z, vb.w};
            if (kb + kFlatThreads * 4 < k_end) load4(kb + kFlatThreads * 4 + grp);
            // records are sorted by A-read: in most groups one lookup of the read's base serves all four
            const int base0 = a[0] >= 0 ? __ldg(F.rbase + a[0]) : -1;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                int base = base0;
                if (a[i] != a[0]) base = a[i] >= 0 ? __ldg(F.rbase + a[i]) : -1;
                bool valid = base >= 0;
                if (self_records) valid = valid && b[i] != a[i];  // filter.cpp:538-547
                if (valid) scatter_record(hist, base, s[i], e[i], C);
            }
        }
    }
    __syncthreads();

    // ---- one prefix sum over the whole batch + the two bit maps.  zero <=> high half <= MIN_COV
    // <=> the word, as a signed integer, is below (MIN_COV + 1) << 16 (the low half is >= 0).
    const int zthr = (MIN_COV + 1 > 32767 ? 32767 : MIN_COV + 1) << 16;
    const int RJ = min(P.min_repeat_annotation_threshold, P.max_repeat_annotation_threshold);
    uint32_t carry = 0;
    for (int pass = 0; pass < npass; pass++) {
        const int j0 = pass * kFlatPass + tid * kFlatItems;
        const int sx = sw(j0) ^ j0;  // the flipped bits: common to the whole 16-word chunk
        uint32_t v[kFlatItems];
        uint32_t cbits = 0;
        if (j0 < nb) {
#pragma unroll
            for (int q = 0; q < kFlatItems / 4; q++) {
                const uint4 x = lds128(hist + ((j0 + 4 * q) ^ sx));
                v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
            }
#pragma unroll
            for (int i = 0; i < kFlatItems; i++) {
                const int g = (int)(int16_t)(v[i] & 0xffffu);  // cov0[j0 + i] - cov0[j0 + i - 1]
                cbits |= (g > RJ || g < -RJ) ? (1u << i) : 0u;
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
        uint32_t zbits = 0;
        if (j0 < nb) {
#pragma unroll
            for (int i = 0; i < kFlatItems; i++) {
                v[i] += pre;
                zbits |= ((int)v[i] < zthr) ? (1u << i) : 0u;
            }
#pragma unroll
            for (int q = 0; q < kFlatItems / 4; q++)
                sts128(hist + ((j0 + 4 * q) ^ sx), make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
        }
        zmap16[pass * kFlatThreads + tid] = (uint16_t)zbits;
        cmap16[pass * kFlatThreads + tid] = (uint16_t)cbits;
    }
    __syncthreads();

    // ---- per read
    const int NHR = P.no_hinge_region;
    const int MINT = P.min_repeat_annotation_threshold, MAXT = P.max_repeat_annotation_threshold;
    for (int read = f0 + tid; read < f1; read += kFlatThreads) {
        const int base = F.rbase[read];
        const int64_t nrec = rv.read_off[read + 1] - rv.read_off[read];
        if (base < 0 || nb == 0 || nrec > Packed<uint32_t>::kMaxCount) {
            out.big_list[atomicAdd(&out.counters[3], 1)] = read;
            continue;
        }
        const int nbz = bins_needed(rd.rlen[read], P);
        const int L0 = F.cov_maxbin[read] + 1;  // length of the cut-off-free profile
        auto H = [&](int j) { return hist[sw(base + j)]; };  // packed coverage of the read's bin j

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
                            if (wr) out.anno_pool[off + n] = make_int2((int)(cur >> 2), (int)(cur & 3u) - 1);
                            n++;
                            cur = nxt;
                        }
                    }
                }
            }
            if (have) {
                if (wr) out.anno_pool[off + n] = make_int2((int)(cur >> 2), (int)(cur & 3u) - 1);
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

        out.mask[read] = mk;
        out.cmask[read] = make_int2(msc, mec);
        out.rflags[read] = flags | (skip_hinges ? kFlagSkipHinge : 0);
        out.anno_ref[read] = make_int2(off, kept);
        if (kept > 0 && !skip_hinges && off >= 0) out.work_list[atomicAdd(&out.counters[1], 1)] = read;
    }

    // ---- optional dump of the cut-off-free profiles for .coverage.txt (filter.cpp:599-602)
    if (DUMP && nb > 0) {
        for (int read = f0 + warp; read < f1; read += kFlatThreads / 32) {
            const int base = F.rbase[read];
            if (base < 0 || rv.read_off[read + 1] - rv.read_off[read] > Packed<uint32_t>::kMaxCount) continue;
            const int L0 = F.cov_maxbin[read] + 1;
            int* dst = out.cov0 + out.cov0_off[read];
            for (int j = lane; j < L0; j += 32) dst[j] = f_lo(hist[sw(base + j)]);
        }
    }
}

// Greedy packing of the reads [lo, hi) into batches of at most kFlatBins histogram words.
//   batch      (first read, words in use) per batch, closed by (hi, 0)
//   rbase      per read: first word of its profile inside its batch; -1 = generic path
//   read_batch per read: its batch (K1 flags the batches that hold A == B records)
void flat_plan(const int* rlen, int lo, int hi, int n_read, int cut_off, std::vector<int2>* batch,
               std::vector<int>* rbase, std::vector<int>* read_batch) {
    batch->clear();
    rbase->assign((size_t)n_read, -1);
    read_batch->assign((size_t)n_read, 0);
    int used = 0;
    for (int r = lo; r < hi; r++) {
        const int nbz = bins_needed(rlen[r], cut_off);
        const bool fits = nbz <= kFlatBins;  // otherwise generic path; still belongs to a batch, which reports it
        if (batch->empty() || (fits && used + nbz > kFlatBins)) {
            batch->push_back(make_int2(r, 0));
            used = 0;
        }
        (*read_batch)[r] = (int)batch->size() - 1;
        if (!fits) continue;
        (*rbase)[r] = used;
        used += nbz;
        batch->back().y = used;
    }
    if (batch->empty()) batch->push_back(make_int2(lo, 0));
    batch->push_back(make_int2(hi, 0));
}

template <int SPREAD>
static void launch_flat(int grid, bool dump, const RecView& rv, const ReadView& rd, const hg_filter_params& P,
                        const FlatParams& F, const MaskAnnoOut& out, cudaStream_t st) {
    if (dump)
        k_mask_anno_flat<SPREAD, true><<<grid, kFlatThreads, 0, st>>>(rv, rd, P, F, out);
    else
        k_mask_anno_flat<SPREAD, false><<<grid, kFlatThreads, 0, st>>>(rv, rd, P, F, out);
}

void launch_mask_anno_flat(const RecView& rv, const ReadView& rd, const hg_filter_params& P,
                           FilterScratch& s, const MaskAnnoOut& out, cudaStream_t st) {
    FlatParams F;
    F.batch = s.flat_batch;
    F.rbase = s.flat_rbase;
    F.cov_maxbin = s.cov_maxbin;
    F.batch_self = s.flat_batch_self;
    F.scal = s.scal;
    const int grid = s.flat_nbatch;
    if (grid <= 0) return;
    g_launches += 1;
    const bool dump = out.cov0 != nullptr;
    switch (s.flat_spread) {
        case 1: launch_flat<1>(grid, dump, rv, rd, P, F, out, st); break;
        case 4: launch_flat<4>(grid, dump, rv, rd, P, F, out, st); break;
        case 16: launch_flat<16>(grid, dump, rv, rd, P, F, out, st); break;
        default: launch_flat<8>(grid, dump, rv, rd, P, F, out, st); break;
    }
}

}  // namespace hg
