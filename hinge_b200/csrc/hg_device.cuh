// Device-side views and small helpers shared by the kernels (sm_100a).
#ifndef HG_DEVICE_CUH
#define HG_DEVICE_CUH
#include <cuda_runtime.h>
#include <stdint.h>

#include "hg_params.h"

namespace hg {

// Overlap records resident in HBM as a struct of arrays (the layout named by
// BASELINE.json north_star) plus the CSR that groups them by A-read.
struct RecView {
    int64_t novl;
    const int32_t* __restrict__ aread;
    const int32_t* __restrict__ bread;
    const int32_t* __restrict__ abpos;
    const int32_t* __restrict__ aepos;
    const int32_t* __restrict__ bbpos;
    const int32_t* __restrict__ bepos;
    const int32_t* __restrict__ flags;
    const int64_t* __restrict__ trace_off;
    const uint8_t* __restrict__ trace;
    int32_t tbytes;
    const int64_t* __restrict__ read_off;  // n_read + 1, global read ids
};

struct ReadView {
    int32_t n_read;
    int32_t r_lo, r_hi;      // reads owned by this context
    const int32_t* __restrict__ rlen;
    const int2* __restrict__ qvmask;  // (start, end) of the longest good-QV run
};

// Resolution of coverage profiles, masks and annotations: hard-coded in the
// reference (filter.cpp:386 `int reso = 40`), a compile-time constant here so
// the divisions become multiplies.
constexpr int kReso = 40;

// bin of an event position on the 40-bp grid: profileCoverage emits entry i
// once every event with pos < i*reso has been consumed
// (/root/reference/src/lib/LAInterface.cpp:4309-4317)
__device__ __forceinline__ int cov_bin(int e, int reso) { return e < 0 ? 0 : e / reso + 1; }

// self_cnt[read] (ingest, k_csr_validate): number of A == B records, plus a flag for reads that
// have a record lying inside one 40-bp bin (abpos / 40 == aepos / 40; never produced by
// daligner, whose alignments are >= 1000 bp, but legal input)
constexpr int kSelfDegenerate = 1 << 30;
__device__ __forceinline__ int self_count(int v) { return v & (kSelfDegenerate - 1); }
__device__ __forceinline__ bool is_degenerate(int v) { return (v & kSelfDegenerate) != 0; }

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane_id() >= d) v += o;
    }
    return v;
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xffffffffu, v, d);
        v = o > v ? o : v;
    }
    return v;
}

// Two signed counters packed in one word: low half = profile without cut-off,
// high half = profile with cut-off.  Sums of packed words stay decodable while
// both halves fit their width, so ONE shared-memory histogram and ONE warp scan
// serve both coverage profiles.
template <typename W>
struct Packed;
template <>
struct Packed<uint32_t> {
    static constexpr int kMaxCount = 32000;  // pile-ups deeper than this use the 64-bit path
    static __device__ __forceinline__ uint32_t one_lo() { return 1u; }
    static __device__ __forceinline__ uint32_t one_hi() { return 1u << 16; }
    static __device__ __forceinline__ int lo(uint32_t v) { return (int)(int16_t)(v & 0xffffu); }
    static __device__ __forceinline__ int hi(uint32_t v) {
        return ((int32_t)(v - (uint32_t)(int32_t)(int16_t)(v & 0xffffu))) >> 16;
    }
    static __device__ __forceinline__ void add(uint32_t* p, uint32_t v) { atomicAdd(p, v); }
};
template <>
struct Packed<unsigned long long> {
    static constexpr int kMaxCount = 2000000000;
    typedef unsigned long long W;
    static __device__ __forceinline__ W one_lo() { return 1ull; }
    static __device__ __forceinline__ W one_hi() { return 1ull << 32; }
    static __device__ __forceinline__ int lo(W v) { return (int)(int32_t)(v & 0xffffffffull); }
    static __device__ __forceinline__ int hi(W v) {
        return (int)(((long long)(v - (W)(long long)(int32_t)(v & 0xffffffffull))) >> 32);
    }
    static __device__ __forceinline__ void add(W* p, W v) { atomicAdd(p, v); }
};

// The masks of all reads as K4 sees them.  Sharded runs exchange them between phase 2 and
// phase 3; every mask bound is a multiple of g = gcd(40, tspace) (40-bp coverage bins, tspace-bp
// QV tiles), so when the longest read has fewer than 65536 such units both bounds travel in
// one 32-bit word (4 B per read on the wire instead of 8).
struct MaskView {
    const int2* full;
    const uint32_t* packed;  // nullptr = use `full`
    int g;
    __device__ __forceinline__ int2 get(int b) const {
        if (!packed) return full[b];
        const uint32_t w = __ldg(packed + b);
        return make_int2((int)(w & 0xffffu) * g, (int)(w >> 16) * g);
    }
};

// ---- phase exchange of a sharded filter run through NVLink peer memory --------------------
// Every rank owns one exchange block (cudaMalloc, mapped into the other ranks through CUDA IPC
// or peer access) and every rank's kernels store straight into all of them:
//   flags      arrival counters, one per source rank and exchange (epoch numbers, monotone)
//   hist       the ranks' parts of the mean-coverage histogram (filter.cpp:642-678), one row each
//   mask_pk    the packed masks of ALL reads (MaskView); K2 stores a read's word into every block
// A consumer kernel spins on its OWN block's flags (local memory) until every rank's epoch has
// arrived; a wall-clock limit turns a lost peer into an error instead of a hung GPU.
constexpr int kMaxPeers = 16;
constexpr int kPeerHistWords = 4100;  // 4096 bins + valid count + padding
constexpr size_t kPeerFlagBytes = 256;
constexpr size_t kPeerHistBytes = sizeof(uint32_t) * kMaxPeers * kPeerHistWords;
constexpr unsigned long long kPeerTimeoutNs = 4000000000ull;

struct PeerView {
    int rank, world;  // world <= 1: no exchange
    unsigned epoch;
    uint8_t* base[kMaxPeers];
    __host__ __device__ uint32_t* flag_hist(int r) const { return reinterpret_cast<uint32_t*>(base[r]); }
    __host__ __device__ uint32_t* flag_mask(int r) const { return reinterpret_cast<uint32_t*>(base[r]) + kMaxPeers; }
    __host__ __device__ uint32_t* flag_ovf(int r) const { return reinterpret_cast<uint32_t*>(base[r]) + 2 * kMaxPeers; }
    __host__ __device__ uint32_t* hist(int r, int src) const {
        return reinterpret_cast<uint32_t*>(base[r] + kPeerFlagBytes) + (size_t)src * kPeerHistWords;
    }
    __host__ __device__ uint32_t* mask_pk(int r) const {
        return reinterpret_cast<uint32_t*>(base[r] + kPeerFlagBytes + kPeerHistBytes);
    }
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned ld_acquire_sys(const uint32_t* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Spins until the flag has reached `epoch` (wrap-around safe); false on time-out.
__device__ __forceinline__ bool peer_wait(const uint32_t* flag, unsigned epoch) {
    const unsigned long long t0 = global_timer_ns();
    while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
        __nanosleep(100);
        if (global_timer_ns() - t0 > kPeerTimeoutNs) return false;
    }
    return true;
}

struct MaskAnnoOut {
    int2* mask;        // .mas
    uint32_t* mask_pk; // optional packed copy (MaskView), unit mask_g
    int mask_g;
    int2* cmask;       // .cmas (bin coordinates)
    uint8_t* rflags;   // kFlag*
    int2* anno_ref;    // (offset into pool, count) per read
    int2* anno_pool;   // (pos, type)
    uint8_t* hinge_keep;  // per pooled annotation: is a hinge (cleared here, set by K4)
    int anno_cap;
    int* counters;     // [0] pool used  [1] work-list length  [2] overflow flag  [3] big-list length
    int4* work_items;  // reads that need hinge calling: 3 x int4 each, see push_work_item
    int* big_list;     // reads whose profile does not fit the shared-memory path
    int* cov0;         // optional dump of the cut-off-free profile (coverage.txt)
    const int64_t* cov0_off;
    PeerView peer;     // world > 1: the packed word also goes into every other rank's array
};

__device__ __forceinline__ void store_mask(const MaskAnnoOut& out, int read, int2 mk) {
    out.mask[read] = mk;
    if (out.mask_pk) {
        const uint32_t w = (uint32_t)(mk.x / out.mask_g) | ((uint32_t)(mk.y / out.mask_g) << 16);
        out.mask_pk[read] = w;
        for (int r = 0; r < out.peer.world; r++)  // posted NVLink writes; k_peer_signal publishes them
            if (r != out.peer.rank) out.peer.mask_pk(r)[read] = w;
    }
}

// K4 is a chain of dependent loads per read; everything it needs to get going travels in the
// work item itself (one 48-byte read instead of read id -> offsets / mask / annotation ref ->
// annotations): (read, pile-up size, first record lo, hi) (mask.x, mask.y, first annotation,
// annotations) (pos0, type0, pos1, type1: the first two annotations, most reads have no more)
__device__ __forceinline__ void write_work_item(const MaskAnnoOut& out, int slot, int read, int64_t o0, int np,
                                                int2 mk, int off, int kept) {
    const int2 a0 = out.anno_pool[off], a1 = kept > 1 ? out.anno_pool[off + 1] : make_int2(0, 0);
    int4* it = out.work_items + 3 * (size_t)slot;
    it[0] = make_int4(read, np, (int)(o0 & 0xffffffffll), (int)(o0 >> 32));
    it[1] = make_int4(mk.x, mk.y, off, kept);
    it[2] = make_int4(a0.x, a0.y, a1.x, a1.y);
}
__device__ __forceinline__ void push_work_item(const MaskAnnoOut& out, int read, int64_t o0, int np, int2 mk,
                                               int off, int kept) {
    write_work_item(out, atomicAdd(&out.counters[1], 1), read, o0, np, mk, off, kept);
}

// Histogram words one read needs: every event bin of both profiles (cut-off of either
// sign) plus two trailing empty bins.
__host__ __device__ __forceinline__ int bins_needed(int rlen, int cut_off) {
    return (rlen + (cut_off < 0 ? -cut_off : cut_off)) / kReso + 3;
}
__device__ __forceinline__ int bins_needed(int rlen, const hg_filter_params& P) {
    return bins_needed(rlen, P.cut_off);
}

}  // namespace hg
#endif
