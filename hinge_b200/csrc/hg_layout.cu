// Kernels of the `hinge maximal` and `hinge layout` stages (sm_100a).
//
//   K5 classify_pairs   per (A,B) pair: top two overlaps by total length, trim to
//                       both reads' masks along the trace, classify
//                       (maximal.cpp:65-134,780-858; hinging.cpp:473-602;
//                        LAInterface.cpp:4552-4683,4721-4781)
//   K5 contain_*        containment recurrence of maximal.cpp:780-858 as a
//                       monotone fixed point over ascending read ids
//   K6 sort_candidates  order-exact weight sort of each read's extension
//                       candidates                     (hinging.cpp:1066-1071)
//   K6 hinge_graph      hinge kill pass + hinge-graph edges through the trace
//                       (hinging.cpp:1262-1321,1365-1640; LAInterface.cpp:4498-4546)
//   K6 best_extension   the best-overlap scoring loop  (hinging.cpp:1911-2148)
#include <algorithm>

#include "hg_device.cuh"
#include "hg_layout.h"
#include "hg_order.h"

namespace hg {

extern int64_t g_launches;

// ------------------------------------------------------------------ classification

struct Match {
    int as, ae, bs, be, comp;     // raw match, B on its forward strand
    int eas, eae, ebs, ebe;       // trimmed to the masks
    int type, weight, length;
    bool active;
};

__device__ __forceinline__ int trace_value(const RecView& rv, int64_t off, int k) {
    if (rv.tbytes == 1) return rv.trace[off + k];
    return reinterpret_cast<const uint16_t*>(rv.trace + off)[k];
}

// ProcessAlignment = trim_overlap + AddTypesAsymmetric
// (maximal.cpp:65-134; LAInterface.cpp:4552-4683, 4721-4781).
//
// trim_overlap walks the trace points (A advances to the next multiple of 100, B by the trace's
// b-delta) and takes the FIRST point inside both effective reads as the trimmed start and the LAST
// one as the trimmed end.  The b-deltas are unsigned, so along the walk both coordinates are
// monotone and the end condition holds for a prefix of the points only; the final point is pushed
// separately (the match's own end).  That allows two early exits without changing any result:
//   * the final point satisfies the end condition -> it is the end; walk only until the start is found
//   * the end condition has turned false          -> the end is known; if the start has not been
//     found by then the match is inactive (start index >= end index), whatever comes later
// About half of all overlaps stick out of a mask at their far end and are walked to (almost) the
// end; the others are done after a few points.
__device__ Match classify_record(const RecView& rv, const int* __restrict__ rlen,
                                 const int2* __restrict__ mask, int64_t k, int a, int b,
                                 const hg_layout_params& P) {
    Match m;
    m.as = __ldg(rv.abpos + k);
    m.ae = __ldg(rv.aepos + k);
    m.comp = __ldg(rv.flags + k) & 1;
    m.bs = __ldg(rv.bbpos + k);
    m.be = __ldg(rv.bepos + k);
    if (m.comp) {
        const int bl = __ldg(rlen + b);
        const int t = bl - m.be;
        m.be = bl - m.bs;
        m.bs = t;
    }
    const int2 EA = mask[a], EB = mask[b];
    m.eas = m.as; m.eae = m.ae; m.ebs = m.bs; m.ebe = m.be;
    const int64_t toff = __ldg(rv.trace_off + k);
    const int tlen = (int)((__ldg(rv.trace_off + k + 1) - toff) / rv.tbytes);
    const int inner = max(tlen / 2 - 1, 0);
    const int npts = inner + 2;
    // Point idx of the walk (LAInterface.cpp:4569-4590): A = abpos for idx 0, then the multiples of
    // 100 above it, (abpos / 100 + idx) * 100; B = its start plus / minus the first idx b-deltas.
    // u = +B (or -B for a complemented match) grows along the walk, so do the A coordinates, and all
    // four mask tests become "index >= / <= a bound known up front" and "u >= / <= a bound":
    //   start (first point inside both effective reads):  idx >= ia_s  and  u >= u_s
    //   end   (last such point):                          idx <= ia_e  and  u <= u_e
    const int sign = 1 - 2 * m.comp;
    const int u_s = m.comp ? -EB.y : EB.x, u_e = m.comp ? -EB.x : EB.y;
    const int a100 = m.as / 100;
    const int ia_s = m.as >= EA.x ? 0 : max(1, (EA.x + 99) / 100 - a100);   // EA.x > as >= 0 here
    const int ia_e = EA.y < m.as ? -1 : (EA.y < 0 ? -1 : max(0, min(inner, EA.y / 100 - a100)));
    // the final point (idx npts - 1): the match's own end
    const int fa = m.ae, fb = m.comp ? m.bs : m.be;
    const int fu = sign * fb;
    const bool end_final = fa <= EA.y && fu <= u_e;
    const int u0 = sign * (m.comp ? m.be : m.bs);
    const uint8_t* __restrict__ t8 = rv.trace + toff;
    const uint16_t* __restrict__ t16 = reinterpret_cast<const uint16_t*>(rv.trace + toff);
    const bool wide = rv.tbytes != 1;
    // Both tests are monotone along the walk (start: false ... false true ... true, end: true ... true
    // false ... false), so the walk only COUNTS: n_s points before the start, n_e points that pass the
    // end test, and u at the first / last such point is a running min / max.  Four trace values are
    // fetched side by side and the exits are tested once per four points (the exits depend on the
    // values: a one-at-a-time loop is a chain of load latencies); what the extra points of the last
    // group may add is taken back below.  About a dozen instructions per point.
    int n_s = 0, n_e = 0, seen = 0;
    int u = u0, u_start = 0x7fffffff, u_end = -0x7fffffff - 1;
    bool done = false;
    for (int idx0 = 0; idx0 <= inner && !done; idx0 += 4) {
        int d[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int idx = idx0 + i;
            d[i] = (idx > 0 && idx <= inner) ? (wide ? (int)t16[2 * idx - 1] : (int)t8[2 * idx - 1]) : 0;
        }
        const int rs = ia_s - idx0, re = ia_e - idx0, rn = inner - idx0;   // the index tests, relative to the group
#pragma unroll
        for (int i = 0; i < 4; i++) {
            u += d[i];
            const bool in = i <= rn;
            const bool s_ok = in && i >= rs && u >= u_s;
            const bool e_ok = in && i <= re && u <= u_e;
            n_s += (in && !s_ok) ? 1 : 0;
            n_e += e_ok ? 1 : 0;
            u_start = min(u_start, s_ok ? u : 0x7fffffff);
            u_end = max(u_end, e_ok ? u : -0x7fffffff - 1);
        }
        seen = min(idx0 + 4, inner + 1);
        // the reference's walk would stop at the first point where this holds
        done = end_final ? n_s < seen : n_e < seen;
    }
    bool have_start = n_s < seen;
    int start_idx = have_start ? n_s : npts;
    // without a final end point the reference's walk stops at the first point that fails the end test
    // (index n_e): a start behind that point was never seen
    if (!end_final && n_e < seen && start_idx > n_e) {
        have_start = false;
        start_idx = npts;
    }
    int end_idx = n_e > 0 ? n_e - 1 : 0;
    if (n_e == 0) u_end = 0;
    if (have_start) {
        m.eas = start_idx == 0 ? m.as : (a100 + start_idx) * 100;
        if (!m.comp) m.ebs = u_start; else m.ebe = -u_start;
    } else if (fa >= EA.x && fu >= u_s) {
        m.eas = fa;
        if (!m.comp) m.ebs = fb; else m.ebe = fb;
        start_idx = npts - 1;
    }
    if (end_final) {
        m.eae = fa;
        if (!m.comp) m.ebe = fb; else m.ebs = fb;
        end_idx = npts - 1;
    } else if (ia_e >= 0 && (end_idx > 0 || (m.as <= EA.y && sign * (m.comp ? m.be : m.bs) <= u_e))) {
        // the last cumulative point that passed (idx 0 included)
        m.eae = end_idx == 0 ? m.as : (a100 + end_idx) * 100;
        if (!m.comp) m.ebe = u_end; else m.ebs = -u_end;
    }
    m.active = a != b && !(start_idx >= end_idx);
    m.type = HG_UNDEFINED;
    if ((m.ebe - m.ebs) < P.aln_threshold || (m.eae - m.eas) < P.aln_threshold || !m.active) {
        m.active = false;
        m.type = HG_NOT_ACTIVE;
    } else {
        const int th = P.theta, th2 = P.theta2;
        const int al = m.eas - EA.x, ar = EA.y - m.eae;
        int bl = m.ebs - EB.x, br = EB.y - m.ebe;
        if (m.comp) {
            const int t = bl;
            bl = br;
            br = t;
        }
        if (max(al, ar) < th && min(bl, br) > th2)
            m.type = HG_BCOVERA;
        else if (max(bl, br) < th && min(al, ar) > th2)
            m.type = HG_ACOVERB;
        else if (min(al, ar) > th)
            m.type = HG_INTERNAL;
        else if (al <= th) {
            if (br <= th && bl >= th)
                m.type = HG_BACKWARD;
            else if (br >= th && bl >= th)
                m.type = HG_BACKWARD_INTERNAL;
        } else if (ar <= th) {
            if (bl <= th && br >= th)
                m.type = HG_FORWARD;
            else if (bl >= th && br >= th)
                m.type = HG_FORWARD_INTERNAL;
            else
                m.type = HG_UNDEFINED;
        }
    }
    m.weight = m.eae - m.eas + m.ebe - m.ebs;
    m.length = m.ae - m.as + m.be - m.bs;
    return m;
}

__device__ __forceinline__ int raw_length(const RecView& rv, int64_t k) {
    // compare_overlap's key (LAInterface.cpp:4884); the B span is strand-invariant
    return (rv.aepos[k] - rv.abpos[k]) + (rv.bepos[k] - rv.bbpos[k]);
}

// ------------------------------------------------------------------ emit (shared)

__device__ void emit_pair(const RecView& rv, const ReadView& rd, const hg_layout_params& P,
                          const int2* __restrict__ mask, const uint8_t* __restrict__ active,
                          int mode, int a, int b, int64_t first, const int64_t top[2],
                          uint8_t* __restrict__ rtype, const PairOut& po) {
    int pair_slot = -1;
    if (mode == 1) {
        pair_slot = atomicAdd(&po.counters[4], 1);
        if (pair_slot < po.pair_cap) {
            po.pairs[pair_slot] = make_int4(a, b, (int)(first & 0x7fffffff), (int)(first >> 31));
        } else {
            atomicExch(&po.counters[3], 1);
        }
    }
    for (int r = 0; r < 2; r++) {
        if (top[r] < 0) break;
        if (r == 1 && !P.use_two_matches) break;
        const Match m = classify_record(rv, rd.rlen, mask, top[r], a, b, P);
        if (mode == 0) {
            rtype[top[r]] = (uint8_t)m.type;
        } else {
            const bool fwd = m.type == HG_FORWARD || m.type == HG_FORWARD_INTERNAL;
            const bool bwd = m.type == HG_BACKWARD || m.type == HG_BACKWARD_INTERNAL;
            if (m.type == HG_BCOVERA && active[b]) po.contained_flag[a] = 1;  // hinging.cpp:598
            if (fwd || bwd) {
                const int slot = atomicAdd(&po.counters[5], 1);
                if (slot < po.cand_cap) {
                    Cand c;
                    c.a = a; c.b = b; c.type = m.type; c.comp = m.comp; c.weight = m.weight;
                    c.length = m.length;
                    c.eas = m.eas; c.eae = m.eae; c.ebs = m.ebs; c.ebe = m.ebe;
                    c.as = m.as; c.ae = m.ae; c.bs = m.bs; c.be = m.be;
                    c.rec = top[r];
                    c.rank = r;
                    c.pad = 0;
                    po.cands[slot] = c;
                } else {
                    atomicExch(&po.counters[3], 1);
                }
            }
        }
    }
}

// One thread per record; the first record of every (A,B) run does the pair.
// mode 0 (maximal): every pair of an active A; marks BCOVERA records.
// mode 1 (layout): pairs whose A and B are both active; appends pair keys and
// the classified top-two records to compact lists.
__global__ void __launch_bounds__(256)
k_classify_pairs(RecView rv, ReadView rd, hg_layout_params P, const int2* __restrict__ mask,
                 const uint8_t* __restrict__ active, int mode, int sort_passes,
                 uint8_t* __restrict__ rtype, PairOut po) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= rv.novl) return;
    const int a = rv.aread[k], b = rv.bread[k];
    if (k > 0 && rv.aread[k - 1] == a && rv.bread[k - 1] == b) return;  // not a pair leader
    if (!active[a]) return;
    if (mode == 1 && !(active[b] && P.keep_only_maximal)) return;
    // run length
    int64_t e = k + 1;
    while (e < rv.novl && rv.aread[e] == a && rv.bread[e] == b) e++;
    const int m = (int)(e - k);
    int64_t top[2] = {k, -1};
    if (m > 16) {
        // std::sort is only stable up to 16 elements: leave it to the order-exact kernel
        const int slot = atomicAdd(&po.counters[0], 1);
        if (slot < po.big_cap) {
            po.big_pairs[slot] = k;
        } else {
            atomicExch(&po.counters[3], 1);
        }
        return;
    }
    if (m > 1) {
        // insertion sort by length, descending, is stable: best = earliest maximum,
        // second = next in (length desc, file order)
        int k1 = raw_length(rv, k), k2 = -2147483647 - 1;
        int64_t i2 = -1;
        for (int64_t i = k + 1; i < e; i++) {
            const int key = raw_length(rv, i);
            if (key > k1) {
                k2 = k1; i2 = top[0];
                k1 = key; top[0] = i;
            } else if (i2 < 0 || key > k2) {
                k2 = key; i2 = i;
            }
        }
        top[1] = i2;
    }
    (void)sort_passes;
    emit_pair(rv, rd, P, mask, active, mode, a, b, k, top, rtype, po);
}

// Pairs with more than 16 records: libstdc++'s introsort on (length, index),
// applied as often as the reference applies it (twice in maximal.cpp:647-654 +
// :791, once in hinging.cpp:534).
__global__ void k_classify_big_pairs(RecView rv, ReadView rd, hg_layout_params P,
                                     const int2* __restrict__ mask,
                                     const uint8_t* __restrict__ active, int mode, int sort_passes,
                                     uint8_t* __restrict__ rtype, PairOut po) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int nbig = min(po.counters[0], po.big_cap);
    if (t >= nbig) return;
    const int64_t k = po.big_pairs[t];
    const int a = rv.aread[k], b = rv.bread[k];
    int64_t e = k + 1;
    while (e < rv.novl && rv.aread[e] == a && rv.bread[e] == b) e++;
    const int m = (int)(e - k);
    const int off = atomicAdd(&po.counters[1], m);
    if (off + m > po.sort_cap) {
        atomicExch(&po.counters[3], 1);
        return;
    }
    KeyIdx2* s = po.sort_scratch + off;
    for (int i = 0; i < m; i++) {
        s[i].key = raw_length(rv, k + i);
        s[i].idx = i;
    }
    for (int p = 0; p < sort_passes; p++) std_sort_exact(s, m, KeyIdx2Greater());
    int64_t top[2] = {k + s[0].idx, k + s[1].idx};
    emit_pair(rv, rd, P, mask, active, mode, a, b, k, top, rtype, po);
}

// ------------------------------------------------------------------ K5, warp per read (maximal)
//
// One warp per active A-read, lanes over its records in file order (coalesced columns).  A lane
// finds its pair's extent (most pairs have one record: both neighbours differ), ranks its record
// inside the pair by (length descending, file order) -- which is what the reference's std::sort
// leaves in front for pairs of at most 16 records (insertion sort: stable) -- and classifies it
// if it is among the top two.  Larger pairs go to the order-exact kernel below.  Every record of
// an active read gets its type written (255 = not among the top two).
__global__ void __launch_bounds__(128)
k_classify_reads(RecView rv, ReadView rd, hg_layout_params P, const int2* __restrict__ mask,
                 const uint8_t* __restrict__ active, uint8_t* __restrict__ rtype, PairOut po) {
    const int lane = lane_id();
    const int a = rd.r_lo + (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (a >= rd.r_hi || !active[a]) return;
    const int64_t o0 = rv.read_off[a], o1 = rv.read_off[a + 1];
    const int ntop = P.use_two_matches ? 2 : 1;
    for (int64_t kb = o0; kb < o1; kb += 32) {
        const int64_t k = kb + lane;
        const bool valid = k < o1;
        const int b = valid ? __ldg(rv.bread + k) : -1;
        int bp = __shfl_up_sync(0xffffffffu, b, 1), bn = __shfl_down_sync(0xffffffffu, b, 1);
        if (lane == 0) bp = kb > o0 ? __ldg(rv.bread + kb - 1) : -2;
        if (lane == 31) bn = kb + 32 < o1 ? __ldg(rv.bread + kb + 32) : -3;
        if (!valid) continue;
        bool take = true;
        if (b == bp || b == bn) {
            int64_t ps = k, pe = k + 1;
            while (ps > o0 && __ldg(rv.bread + ps - 1) == b) ps--;
            while (pe < o1 && __ldg(rv.bread + pe) == b) pe++;
            if (pe - ps > 16) {
                // std::sort is only stable up to 16 elements: leave it to the order-exact kernel
                if (k == ps) {
                    const int slot = atomicAdd(&po.counters[0], 1);
                    if (slot < po.big_cap) po.big_pairs[slot] = k; else atomicExch(&po.counters[3], 1);
                }
                continue;  // rtype written by k_classify_big_pairs (255 where it does not)
            }
            const int key = raw_length(rv, k);
            int rank = 0;
            for (int64_t j = ps; j < pe; j++) {
                if (j == k) continue;
                const int kj = raw_length(rv, j);
                rank += (kj > key || (kj == key && j < k)) ? 1 : 0;
            }
            take = rank < ntop;
        }
        uint8_t t = HG_NOT_CLASSIFIED;
        if (take) t = (uint8_t)classify_record(rv, rd.rlen, mask, k, a, b, P).type;
        rtype[k] = t;
    }
}

// Records of pairs with more than 16 records: 255 unless the order-exact kernel classifies them.
__global__ void k_clear_big_pair_types(RecView rv, uint8_t* __restrict__ rtype, PairOut po) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int nbig = min(po.counters[0], po.big_cap);
    if (t >= nbig) return;
    const int64_t k = po.big_pairs[t];
    const int a = rv.aread[k], b = rv.bread[k];
    for (int64_t e = k; e < rv.novl && rv.aread[e] == a && rv.bread[e] == b; e++) rtype[e] = HG_NOT_CLASSIFIED;
}

// ------------------------------------------------------------------ K5, warp per read (layout)
//
// Pairs between maximal reads and their classified top-two overlaps, device-resident and in the
// reference's order: one warp per active A-read, lanes over its records.  Pair leaders whose B is
// active get consecutive pair slots (ballot prefix: file order = ascending B, the insertion order of
// the reference's idx_ab[A], hinging.cpp:477-480), the leader classifies the pair's top two and the
// FORWARD / BACKWARD(_INTERNAL) ones become candidates in consecutive slots of the read's block.

// pairs per read: first pass of the same loop, so the blocks can be laid out without guessing
__global__ void __launch_bounds__(128)
k_layout_count_pairs(RecView rv, ReadView rd, const uint8_t* __restrict__ active, int2* __restrict__ pair_ref,
                     unsigned long long* __restrict__ total) {
    const int lane = lane_id();
    const int a = rd.r_lo + (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (a >= rd.r_hi) return;
    int np = 0;
    if (active[a]) {
        const int64_t o0 = rv.read_off[a], o1 = rv.read_off[a + 1];
        for (int64_t kb = o0; kb < o1; kb += 32) {
            const int64_t k = kb + lane;
            const int b = k < o1 ? __ldg(rv.bread + k) : -1;
            int bp = __shfl_up_sync(0xffffffffu, b, 1);
            if (lane == 0) bp = kb > o0 ? __ldg(rv.bread + kb - 1) : -2;
            const bool lead = k < o1 && b != bp && active[b];
            np += __popc(__ballot_sync(0xffffffffu, lead));
        }
    }
    if (lane == 0) {
        pair_ref[a] = make_int2(0, np);
        if (np) atomicAdd(total, (unsigned long long)np);
    }
}

__device__ __forceinline__ int hash_buckets_for(int n, const int* __restrict__ grow_at,
                                                const int* __restrict__ grow_bkt, int ngrow) {
    int nb = 1;
    for (int i = 0; i < ngrow && grow_at[i] <= n; i++) nb = grow_bkt[i];
    return nb;
}

__device__ __forceinline__ void store_cand(Cand* dst, const Match& m, int a, int b, int64_t rec, int rank) {
    Cand c;
    c.a = a; c.b = b; c.type = m.type; c.comp = m.comp; c.weight = m.weight; c.length = m.length;
    c.eas = m.eas; c.eae = m.eae; c.ebs = m.ebs; c.ebe = m.ebe;
    c.as = m.as; c.ae = m.ae; c.bs = m.bs; c.be = m.be;
    c.rec = rec;
    c.rank = rank;
    c.pad = 0;
    *dst = c;
}

__global__ void __launch_bounds__(128)
k_layout_pairs(RecView rv, ReadView rd, hg_layout_params P, const int2* __restrict__ mask,
               const uint8_t* __restrict__ active, LayoutLists L) {
    const int lane = lane_id();
    const unsigned lt = (1u << lane) - 1u;
    const int a = rd.r_lo + (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (a >= rd.r_hi) return;
    const int np = L.pair_ref[a].y;
    if (np == 0) {
        if (lane == 0) {
            L.cand_ref[a] = make_int2(0, 0);
            L.bkt_ref[a] = 0;
        }
        return;
    }
    int poff = 0, coff = 0, boff = 0;
    if (lane == 0) {
        poff = atomicAdd(&L.counters[0], np);
        coff = atomicAdd(&L.counters[1], 2 * np);
        boff = atomicAdd(&L.counters[2], hash_buckets_for(np, L.grow_at, L.grow_bkt, L.ngrow));
    }
    poff = __shfl_sync(0xffffffffu, poff, 0);
    coff = __shfl_sync(0xffffffffu, coff, 0);
    boff = __shfl_sync(0xffffffffu, boff, 0);
    const int64_t o0 = rv.read_off[a], o1 = rv.read_off[a + 1];
    int pi = 0, ci = 0;
    for (int64_t kb = o0; kb < o1; kb += 32) {
        const int64_t k = kb + lane;
        const int b = k < o1 ? __ldg(rv.bread + k) : -1;
        int bp = __shfl_up_sync(0xffffffffu, b, 1);
        if (lane == 0) bp = kb > o0 ? __ldg(rv.bread + kb - 1) : -2;
        const bool lead = k < o1 && b != bp && active[b];
        const unsigned lm = __ballot_sync(0xffffffffu, lead);
        Match m0, m1;
        int64_t top[2] = {k, -1};
        int v0 = 0, v1 = 0;
        if (lead) {
            int64_t e = k + 1;
            while (e < o1 && __ldg(rv.bread + e) == b) e++;
            const int m = (int)(e - k);
            if (m > 16) {
                // std::sort is only stable up to 16 elements: libstdc++'s introsort on (length, index),
                // once (hinging.cpp:534), in a slice of the sort scratch
                const int off = atomicAdd(&L.counters[4], m);
                if (off + m > L.sort_cap) {
                    atomicExch(&L.counters[3], 1);
                } else {
                    KeyIdx2* sc = L.sort_scratch + off;
                    for (int i = 0; i < m; i++) {
                        sc[i].key = raw_length(rv, k + i);
                        sc[i].idx = i;
                    }
                    std_sort_exact(sc, m, KeyIdx2Greater());
                    top[0] = k + sc[0].idx;
                    top[1] = k + sc[1].idx;
                }
            } else if (m > 1) {
                // insertion sort by length, descending, is stable: best = earliest maximum,
                // second = next in (length desc, file order)
                int k1 = raw_length(rv, k), k2 = -2147483647 - 1;
                int64_t i2 = -1;
                for (int64_t i = k + 1; i < e; i++) {
                    const int key = raw_length(rv, i);
                    if (key > k1) {
                        k2 = k1; i2 = top[0];
                        k1 = key; top[0] = i;
                    } else if (i2 < 0 || key > k2) {
                        k2 = key; i2 = i;
                    }
                }
                top[1] = i2;
            }
            m0 = classify_record(rv, rd.rlen, mask, top[0], a, b, P);
            v0 = (m0.type == HG_FORWARD || m0.type == HG_FORWARD_INTERNAL || m0.type == HG_BACKWARD ||
                  m0.type == HG_BACKWARD_INTERNAL) ? 1 : 0;
            bool bcov = m0.type == HG_BCOVERA;
            if (top[1] >= 0 && P.use_two_matches) {
                m1 = classify_record(rv, rd.rlen, mask, top[1], a, b, P);
                v1 = (m1.type == HG_FORWARD || m1.type == HG_FORWARD_INTERNAL || m1.type == HG_BACKWARD ||
                      m1.type == HG_BACKWARD_INTERNAL) ? 1 : 0;
                bcov = bcov || m1.type == HG_BCOVERA;
            }
            if (bcov) L.contained_flag[a] = 1;  // hinging.cpp:598 (B is active here)
        }
        // consecutive candidate slots in lane (= pair) order
        const int nc = v0 + v1;
        int incl = nc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lead) {
            const int start = ci + incl - nc;
            if (v0) store_cand(L.cands + coff + start, m0, a, b, top[0], 0);
            if (v1) store_cand(L.cands + coff + start + v0, m1, a, b, top[1], 1);
            L.pairs[poff + pi + __popc(lm & lt)] = make_int2(b, (start << 2) | nc);
        }
        pi += __popc(lm);
        ci += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) {
        L.pair_ref[a] = make_int2(poff, np);
        L.cand_ref[a] = make_int2(coff, ci);
        L.bkt_ref[a] = boff;
    }
}

// Pre-sort order + weight sort of every read's candidate lists, one thread per read:
// the reference fills matches_forward / matches_backward while iterating idx_ab[A], a
// std::unordered_map keyed by B (hinging.cpp:532-592), and then std::sorts both by weight
// (hinging.cpp:1066-1071; unstable: ties keep an order that depends on the input order).  The
// iteration order is replayed by hash_iteration_order (hg_order.h), the sort by std_sort_exact.
__global__ void __launch_bounds__(128)
k_order_candidates(LayoutLists L, SelectIO io) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= io.n_read) return;
    int4 rg = make_int4(0, 0, 0, 0);
    const int2 pr = L.pair_ref[a];
    if (io.active[a] && pr.y > 0) {
        const int2 cr = L.cand_ref[a];
        const int2* pairs = L.pairs + pr.x;
        int* out = L.hash_out + pr.x;
        hash_iteration_order([&](int i) { return pairs[i].x; }, pr.y, L.grow_at, L.grow_bkt, L.ngrow,
                             L.hash_next + pr.x, L.hash_bkt + L.bkt_ref[a], out);
        int n = 0;
        for (int half = 0; half < 2; half++) {
            const int lo = n;
            for (int t = 0; t < pr.y; t++) {
                const int info = pairs[out[t]].y;
                for (int r = 0; r < (info & 3); r++) {
                    const int ci = cr.x + (info >> 2) + r;
                    const int ty = io.cands[ci].type;
                    const bool fwd = ty == HG_FORWARD || ty == HG_FORWARD_INTERNAL;
                    if (fwd == (half == 0)) io.order[cr.x + n++] = ci;
                }
            }
            if (half == 0) {
                rg.x = cr.x + lo;
                rg.y = cr.x + n;
            } else {
                rg.z = cr.x + lo;
                rg.w = cr.x + n;
            }
            const int cnt = n - lo;
            if (cnt > 1) {
                KeyIdx2* s = io.sort_scratch + cr.x + lo;
                for (int t = 0; t < cnt; t++) {
                    s[t].idx = io.order[cr.x + lo + t];
                    s[t].key = io.cands[s[t].idx].weight;
                }
                std_sort_exact(s, cnt, KeyIdx2Greater());
                for (int t = 0; t < cnt; t++) io.order[cr.x + lo + t] = s[t].idx;
            }
        }
    }
    io.ranges_out[a] = rg;
}

// active &= !contained ("[contained] Should not happen", hinging.cpp:598-601); counts them
__global__ void k_apply_contained(int n, const uint8_t* __restrict__ contained, uint8_t* __restrict__ active,
                                  int* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && contained[i] && active[i]) {
        active[i] = 0;
        atomicAdd(count, 1);
    }
}

// the chosen candidates, compacted for the host: (candidate, hinge_pos) per read and direction
__global__ void k_gather_chosen(int n_read, const int2* __restrict__ chosen, const Cand* __restrict__ cands,
                                Cand* __restrict__ out, int2* __restrict__ out_ref, int* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n_read) return;
    const int2 ch = chosen[i];
    if (ch.x < 0) return;
    const int slot = atomicAdd(count, 1);
    out[slot] = cands[ch.x];
    out_ref[slot] = make_int2(i, ch.y);
}

// ------------------------------------------------------------------ containment

// state: 0 unknown, 1 survives (maximal), 2 removed
__global__ void k_contain_init(RecView rv, ReadView rd, const uint8_t* __restrict__ active0,
                               const uint8_t* __restrict__ rtype, uint8_t* __restrict__ state) {
    const int i = rd.r_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rd.r_hi) return;
    if (!active0[i]) {
        state[i] = 2;
        return;
    }
    bool hi = false, lo = false;
    for (int64_t k = rv.read_off[i]; k < rv.read_off[i + 1]; k++) {
        if (rtype[k] != HG_BCOVERA) continue;
        const int b = rv.bread[k];
        if (!active0[b]) continue;
        // B > A is still active when A is processed (maximal.cpp:809: reads[B]->active)
        if (b > i) hi = true; else lo = true;
    }
    state[i] = hi ? 2 : (lo ? 0 : 1);
}

__global__ void k_contain_step(RecView rv, ReadView rd, const uint8_t* __restrict__ active0,
                               const uint8_t* __restrict__ rtype, volatile uint8_t* state,
                               int* __restrict__ remaining) {
    const int i = rd.r_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rd.r_hi || state[i] != 0) return;
    bool any_alive = false, any_unknown = false;
    for (int64_t k = rv.read_off[i]; k < rv.read_off[i + 1]; k++) {
        if (rtype[k] != HG_BCOVERA) continue;
        const int b = rv.bread[k];
        if (b >= i || !active0[b]) continue;
        const uint8_t s = state[b];  // B < A: B's FINAL state decides (maximal.cpp:853-854)
        any_alive = any_alive || s == 1;
        any_unknown = any_unknown || s == 0;
    }
    if (any_alive)
        state[i] = 2;
    else if (!any_unknown)
        state[i] = 1;
    else
        atomicAdd(remaining, 1);
}

// The same recurrence on compact lists.  Containment (maximal.cpp:780-858) only ever looks at
// containing reads: a read with an active container of HIGHER id is removed whatever happens
// (that container has not been visited when the reference visits the read), one without
// containers survives, and the rest -- "unknown": all containers have lower ids -- depend on the
// final state of those.  k_contain_lists settles the first two groups and writes, for every
// unknown read, the list of its lower-id containers; k_contain_resolve then iterates over the
// unknown reads only (a few per cent of the reads, a handful of list entries each) inside ONE CTA,
// no host round trips.  Sharded runs gather the lists and states of all ranks before the resolve
// (hg_maximal_phase1 / hg_maximal_phase2).
__global__ void __launch_bounds__(128)
k_contain_lists(RecView rv, ReadView rd, ContainIO io) {
    const int lane = lane_id();
    const unsigned lt = (1u << lane) - 1u;
    const int i = rd.r_lo + (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (i >= rd.r_hi) return;
    if (!io.active0[i]) {
        if (lane == 0) io.state[i] = 2;
        return;
    }
    const int64_t o0 = rv.read_off[i], o1 = rv.read_off[i + 1];
    bool hi = false;
    int nlo = 0;
    for (int64_t kb = o0; kb < o1; kb += 32) {
        const int64_t k = kb + lane;
        bool lo = false;
        if (k < o1 && io.rtype[k] == HG_BCOVERA) {
            const int b = __ldg(rv.bread + k);
            if (io.active0[b]) {
                // B > A is still active when A is processed (maximal.cpp:809: reads[B]->active)
                if (b > i) hi = true; else lo = true;
            }
        }
        nlo += __popc(__ballot_sync(0xffffffffu, lo));
    }
    hi = __any_sync(0xffffffffu, hi);
    if (hi || nlo == 0) {
        if (lane == 0) io.state[i] = hi ? 2 : 1;
        return;
    }
    int slot = 0, off = 0;
    if (lane == 0) {
        slot = atomicAdd(&io.counters[0], 1);
        off = atomicAdd(&io.counters[1], nlo);
        if (slot >= io.unk_cap || off + nlo > io.pool_cap) {
            atomicExch(&io.counters[2], 1);
            slot = -1;
        } else {
            io.unk[slot] = make_int4(i, off, nlo, 0);
        }
        io.state[i] = 0;
    }
    slot = __shfl_sync(0xffffffffu, slot, 0);
    off = __shfl_sync(0xffffffffu, off, 0);
    if (slot < 0) return;
    int n = 0;
    for (int64_t kb = o0; kb < o1; kb += 32) {
        const int64_t k = kb + lane;
        bool lo = false;
        int b = 0;
        if (k < o1 && io.rtype[k] == HG_BCOVERA) {
            b = __ldg(rv.bread + k);
            lo = io.active0[b] && b < i;
        }
        const unsigned m = __ballot_sync(0xffffffffu, lo);
        if (lo) io.pool[off + n + __popc(m & lt)] = b;
        n += __popc(m);
    }
}

// One sweep over the unknown reads with the whole grid: settles every read whose containers are all
// settled.  Most chains are one or two links long, so a few of these leave only a handful of reads to
// the single-CTA loop below.
__global__ void __launch_bounds__(256)
k_contain_sweep(const int4* __restrict__ unk, const int* __restrict__ counts, int world, int unk_stride,
                const int* __restrict__ pool, int pool_stride, volatile uint8_t* state) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (r >= world || e >= min(counts[2 * r], unk_stride)) return;
    const int4 u = unk[(size_t)r * unk_stride + e];
    if (state[u.x] != 0) return;
    const int* lst = pool + (size_t)r * pool_stride + u.y;
    bool alive = false, unknown = false;
    for (int c = 0; c < u.z; c++) {
        const uint8_t s = state[lst[c]];
        alive = alive || s == 1;
        unknown = unknown || s == 0;
    }
    if (alive)
        state[u.x] = 2;
    else if (!unknown)
        state[u.x] = 1;
}

// unk / pool: `world` segments of unk_stride / pool_stride entries, counts[2 r] = unknown reads of
// segment r.  sweeps_out (may be null) gets the number of sweeps it took.
__global__ void __launch_bounds__(1024)
k_contain_resolve(const int4* __restrict__ unk, const int* __restrict__ counts, int world, int unk_stride,
                  const int* __restrict__ pool, int pool_stride, volatile uint8_t* state, int* sweeps_out) {
    __shared__ int remaining;
    int sweeps = 0;
    for (;;) {
        if (threadIdx.x == 0) remaining = 0;
        __syncthreads();
        for (int r = 0; r < world; r++) {
            const int n = min(counts[2 * r], unk_stride);
            for (int e = threadIdx.x; e < n; e += blockDim.x) {
                const int4 u = unk[(size_t)r * unk_stride + e];
                if (state[u.x] != 0) continue;
                const int* lst = pool + (size_t)r * pool_stride + u.y;
                bool alive = false, unknown = false;
                for (int c = 0; c < u.z; c++) {
                    const uint8_t s = state[lst[c]];  // B < A: B's FINAL state decides (maximal.cpp:853-854)
                    alive = alive || s == 1;
                    unknown = unknown || s == 0;
                }
                if (alive)
                    state[u.x] = 2;
                else if (!unknown)
                    state[u.x] = 1;
                else
                    atomicAdd(&remaining, 1);
            }
        }
        sweeps++;
        __syncthreads();
        const int rem = remaining;
        __syncthreads();
        if (rem == 0 || sweeps > 100000000) break;
    }
    if (threadIdx.x == 0 && sweeps_out) *sweeps_out = sweeps;
}

// ------------------------------------------------------------------ K6: selection

// hinging.cpp:1066-1071: std::sort of each read's candidate lists by weight, descending.
__global__ void k_sort_candidates(SelectIO io) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= io.n_read || !io.active[i]) return;
    const int4 r = io.ranges[i];
    for (int half = 0; half < 2; half++) {
        const int lo = half ? r.z : r.x, hi = half ? r.w : r.y;
        const int n = hi - lo;
        if (n <= 1) continue;
        KeyIdx2* s = io.sort_scratch + lo;
        for (int t = 0; t < n; t++) {
            s[t].idx = io.order[lo + t];
            s[t].key = io.cands[s[t].idx].weight;
        }
        std_sort_exact(s, n, KeyIdx2Greater());
        for (int t = 0; t < n; t++) io.order[lo + t] = s[t].idx;
    }
}

// LOverlap::GetMatchingPosition (LAInterface.cpp:4498-4546)
__device__ int matching_position(const RecView& rv, const Cand& c, int pos_a) {
    if (pos_a < c.as || pos_a > c.ae) return -1;
    const int sign = 1 - 2 * c.comp;
    int cur_a = c.as;
    int cur_b = c.comp ? c.be : c.bs;
    const int64_t toff = rv.trace_off[c.rec];
    const int tlen = (int)((rv.trace_off[c.rec + 1] - toff) / rv.tbytes);
    for (int j = 0; j < tlen / 2 - 1; j++) {
        const int next_a = (cur_a / 100 + 1) * 100;
        if (next_a >= pos_a) return cur_b + pos_a - cur_a;
        cur_b += sign * trace_value(rv, toff, 2 * j + 1);
        cur_a = next_a;
    }
    if (cur_a < pos_a) return cur_b + pos_a - cur_a;
    return -2;
}

// Kill pass (hinging.cpp:1262-1321) and hinge graph (hinging.cpp:1365-1640),
// one thread per active read that carries hinges.
__global__ void k_hinge_graph(RecView rv, hg_layout_params P, SelectIO io) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= io.n_read || !io.active[i]) return;
    const int64_t h0 = io.hv.off[i], h1 = io.hv.off[i + 1];
    if (h0 == h1) return;
    const int4 r = io.ranges[i];
    // kill pass
    for (int half = 0; half < 2; half++) {
        const int lo = half ? r.z : r.x, hi = half ? r.w : r.y;
        for (int t = lo; t < hi; t++) {
            const Cand& m = io.cands[io.order[t]];
            if (!io.active[m.b]) continue;
            for (int64_t k = h0; k < h1; k++) {
                const int pos = io.hv.pos[k], type = io.hv.type[k];
                bool kill;
                if (!half)
                    kill = type == 1 &&
                           ((m.eas < pos + P.kill_hinge_internal && m.type == HG_FORWARD_INTERNAL) ||
                            (m.eas < pos - P.kill_hinge_overlap && m.type == HG_FORWARD));
                else
                    kill = type == -1 &&
                           ((m.eae > pos - P.kill_hinge_internal && m.type == HG_BACKWARD_INTERNAL) ||
                            (m.eae > pos + P.kill_hinge_overlap && m.type == HG_BACKWARD));
                if (kill) io.hinge_alive[k] = 0;
            }
        }
    }
    // hinge graph
    int seq = 0;
    const int S = P.matching_hinge_slack;
    for (int64_t k = h0; k < h1; k++) {
        const int hpos = io.hv.pos[k], htype = io.hv.type[k];
        for (int half = 0; half < 2; half++) {
            const int lo = half ? r.z : r.x, hi = half ? r.w : r.y;
            const int own_type = half ? -1 : 1;
            for (int t = lo; t < hi; t++) {
                const Cand& m = io.cands[io.order[t]];
                if (!io.active[m.b]) continue;
                const int pos_b = matching_position(rv, m, hpos);
                const int req = m.comp ? -htype : htype;
                const int rev = m.comp ? 1 : 0;
                const int b = m.b;
                for (int64_t l = io.hv.off[b]; l < io.hv.off[b + 1]; l++) {
                    const int p2 = io.hv.pos[l];
                    if (p2 < pos_b + S && p2 > pos_b - S && req == io.hv.type[l]) {
                        const int slot = atomicAdd(&io.counters[0], 1);
                        if (slot < io.graph_cap) {
                            GraphRec g;
                            g.owner = i; g.seq = seq; g.flag = 1; g.rev = rev;
                            g.u = (int)k; g.v = (int)l;
                            if (htype == own_type) {
                                g.f[0] = i; g.f[1] = b; g.f[2] = hpos; g.f[3] = p2;
                            } else {
                                g.f[0] = b; g.f[1] = i; g.f[2] = p2; g.f[3] = hpos;
                            }
                            io.graph[slot] = g;
                        } else {
                            atomicExch(&io.counters[3], 1);
                        }
                        seq++;
                    }
                }
                for (int64_t l = io.kv.off[b]; l < io.kv.off[b + 1]; l++) {
                    const int p2 = io.kv.pos[l];
                    if (p2 < pos_b + S && p2 > pos_b - S) {
                        const bool tmatch = req == io.kv.type[l];
                        if (tmatch) {
                            const int slot = atomicAdd(&io.counters[0], 1);
                            if (slot < io.graph_cap) {
                                GraphRec g;
                                g.owner = i; g.seq = seq; g.flag = 0; g.rev = rev; g.u = g.v = -1;
                                if (htype == own_type) {
                                    g.f[0] = i; g.f[1] = b; g.f[2] = hpos; g.f[3] = p2;
                                } else {
                                    g.f[0] = b; g.f[1] = i; g.f[2] = p2; g.f[3] = hpos;
                                }
                                io.graph[slot] = g;
                            } else {
                                atomicExch(&io.counters[3], 1);
                            }
                            seq++;
                        }
                        // forward: inside the type test (hinging.cpp:1472); backward: outside (:1616)
                        const bool push = half ? m.type == HG_BACKWARD : (tmatch && m.type == HG_FORWARD);
                        if (push) {
                            const int slot = atomicAdd(&io.counters[1], 1);
                            if (slot < io.nk_cap) {
                                NkRec n;
                                n.owner = i; n.seq = seq; n.pos = hpos; n.type = htype;
                                io.nkout[slot] = n;
                            } else {
                                atomicExch(&io.counters[3], 1);
                            }
                            seq++;
                        }
                    }
                }
            }
        }
    }
}

// The best-overlap scoring loop (hinging.cpp:1911-2148): per active read, walk
// the candidates in weight order; the first FORWARD not poisoned by a newly
// killed hinge wins unless a FORWARD_INTERNAL that lands on an active hinge of B
// is at most 2 * hinge_slack lighter; mirrored for the backward direction.
__global__ void k_best_extension(hg_layout_params P, SelectIO io) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= io.n_read) return;
    io.chosen[2 * i] = make_int2(-1, -1);
    io.chosen[2 * i + 1] = make_int2(-1, -1);
    if (!io.active[i]) return;
    const int4 r = io.ranges[i];
    const int64_t n0 = io.nk.off[i], n1 = io.nk.off[i + 1];
    int seq = 0;
    for (int half = 0; half < 2; half++) {
        const int lo = half ? r.z : r.x, hi = half ? r.w : r.y;
        const int plain = half ? HG_BACKWARD : HG_FORWARD;
        const int internal = half ? HG_BACKWARD_INTERNAL : HG_FORWARD_INTERNAL;
        int got = 0, got_internal = 0, chosen = -1, hinge_pos = -1, chosen_weight = 0;
        for (int t = lo; t < hi; t++) {
            const int ci = io.order[t];
            const Cand& m = io.cands[ci];
            if (!io.active[m.b]) continue;
            if (m.type == plain && got == 0) {
                bool poisoned = false;
                for (int64_t k = n0; k < n1; k++) {
                    const int kp = io.nk.pos[k], kt = io.nk.type[k];
                    bool hit;
                    if (!half)
                        hit = (m.comp != 1 && kt == -1 && kp > m.ebe) || (m.comp == 1 && kt == 1 && kp < m.ebs);
                    else
                        hit = (m.comp != 1 && kt == 1 && kp < m.ebs) || (m.comp == 1 && kt == -1 && kp > m.ebe);
                    if (hit) {
                        const int slot = atomicAdd(&io.counters[2], 1);
                        if (slot < io.skip_cap) {
                            SkipRec s;
                            s.owner = i; s.seq = seq; s.cand = ci;
                            io.skips[slot] = s;
                        } else {
                            atomicExch(&io.counters[3], 1);
                        }
                        seq++;
                        poisoned = true;
                    }
                }
                if (!poisoned) {
                    chosen = ci;
                    chosen_weight = m.weight;
                    hinge_pos = -1;
                    got = 1;
                }
            } else if (m.type == internal && io.hv.off[m.b + 1] > io.hv.off[m.b] && got_internal == 0) {
                int bpos, want;
                if (!half) {
                    bpos = m.comp == 1 ? m.be : m.bs;
                    want = 1 - 2 * m.comp;
                } else {
                    bpos = m.comp == 1 ? m.bs : m.be;
                    want = -1 + 2 * m.comp;
                }
                for (int64_t k = io.hv.off[m.b]; k < io.hv.off[m.b + 1]; k++) {
                    const int hp = io.hv.pos[k];
                    if (bpos > hp - P.hinge_tolerance && bpos < hp + P.hinge_tolerance &&
                        io.hv.type[k] == want && io.hinge_alive[k]) {
                        if (got == 0 || m.weight > chosen_weight - 2 * P.hinge_slack) {
                            chosen = ci;
                            chosen_weight = m.weight;
                            got = 1;
                            got_internal = 1;
                            hinge_pos = hp;
                        }
                        break;
                    }
                }
            }
        }
        io.chosen[2 * i + half] = make_int2(chosen, hinge_pos);
    }
}

// ------------------------------------------------------------------ launchers

static inline int cdiv(int64_t a, int b) { return (int)((a + b - 1) / b); }

void launch_classify(const RecView& rv, const ReadView& rd, const hg_layout_params& P,
                     const int2* mask, const uint8_t* active, int mode, int sort_passes,
                     uint8_t* rtype, const PairOut& po, cudaStream_t st) {
    cudaMemsetAsync(po.counters, 0, sizeof(int) * 8, st);
    k_classify_pairs<<<cdiv(rv.novl, 256), 256, 0, st>>>(rv, rd, P, mask, active, mode,
                                                         sort_passes, rtype, po);
    // the list length lives on the device; one thread per possible entry
    k_classify_big_pairs<<<cdiv(po.big_cap, 128), 128, 0, st>>>(rv, rd, P, mask, active, mode,
                                                                sort_passes, rtype, po);
    g_launches += 2;
}

void launch_classify_reads(const RecView& rv, const ReadView& rd, const hg_layout_params& P, const int2* mask,
                           const uint8_t* active, int sort_passes, uint8_t* rtype, const PairOut& po,
                           cudaStream_t st) {
    cudaMemsetAsync(po.counters, 0, sizeof(int) * 8, st);
    const int64_t threads = (int64_t)(rd.r_hi - rd.r_lo) * 32;
    k_classify_reads<<<cdiv(threads, 128), 128, 0, st>>>(rv, rd, P, mask, active, rtype, po);
    // the list length lives on the device; one thread per possible entry
    k_clear_big_pair_types<<<cdiv(po.big_cap, 128), 128, 0, st>>>(rv, rtype, po);
    k_classify_big_pairs<<<cdiv(po.big_cap, 128), 128, 0, st>>>(rv, rd, P, mask, active, 0, sort_passes, rtype, po);
    g_launches += 3;
}

void launch_contain_lists(const RecView& rv, const ReadView& rd, const ContainIO& io, cudaStream_t st) {
    cudaMemsetAsync(io.counters, 0, sizeof(int) * 4, st);
    const int64_t threads = (int64_t)(rd.r_hi - rd.r_lo) * 32;
    k_contain_lists<<<cdiv(threads, 128), 128, 0, st>>>(rv, rd, io);
    g_launches += 1;
}

void launch_contain_resolve(const int4* unk, const int* counts, int world, int unk_stride, const int* pool,
                            int pool_stride, uint8_t* state, int* sweeps_out, cudaStream_t st) {
    // racing reads of `state` inside a sweep are harmless: a state only ever goes from 0 to its final
    // value, and a read that still sees 0 just leaves its entry for the next sweep
    const dim3 grid((unsigned)cdiv(std::max(unk_stride, 1), 256), (unsigned)world);
    for (int i = 0; i < 3; i++)
        k_contain_sweep<<<grid, 256, 0, st>>>(unk, counts, world, unk_stride, pool, pool_stride, state);
    k_contain_resolve<<<1, 1024, 0, st>>>(unk, counts, world, unk_stride, pool, pool_stride, state, sweeps_out);
    g_launches += 4;
}

void launch_layout_count_pairs(const RecView& rv, const ReadView& rd, const uint8_t* active, int2* pair_ref,
                               unsigned long long* total, cudaStream_t st) {
    cudaMemsetAsync(total, 0, sizeof(unsigned long long), st);
    k_layout_count_pairs<<<cdiv((int64_t)(rd.r_hi - rd.r_lo) * 32, 128), 128, 0, st>>>(rv, rd, active, pair_ref, total);
    g_launches += 1;
}

void launch_layout_pairs(const RecView& rv, const ReadView& rd, const hg_layout_params& P, const int2* mask,
                         const uint8_t* active, const LayoutLists& L, cudaStream_t st) {
    cudaMemsetAsync(L.counters, 0, sizeof(int) * 8, st);
    k_layout_pairs<<<cdiv((int64_t)(rd.r_hi - rd.r_lo) * 32, 128), 128, 0, st>>>(rv, rd, P, mask, active, L);
    g_launches += 1;
}

void launch_order_candidates(const LayoutLists& L, const SelectIO& io, cudaStream_t st) {
    k_order_candidates<<<cdiv(io.n_read, 128), 128, 0, st>>>(L, io);
    g_launches += 1;
}

void launch_apply_contained(int n, const uint8_t* contained, uint8_t* active, int* count, cudaStream_t st) {
    cudaMemsetAsync(count, 0, sizeof(int), st);
    k_apply_contained<<<cdiv(n, 256), 256, 0, st>>>(n, contained, active, count);
    g_launches += 1;
}

void launch_gather_chosen(int n_read, const int2* chosen, const Cand* cands, Cand* out, int2* out_ref, int* count,
                          cudaStream_t st) {
    cudaMemsetAsync(count, 0, sizeof(int), st);
    k_gather_chosen<<<cdiv(2 * (int64_t)n_read, 256), 256, 0, st>>>(n_read, chosen, cands, out, out_ref, count);
    g_launches += 1;
}

void launch_contain_init(const RecView& rv, const ReadView& rd, const uint8_t* active0,
                         const uint8_t* rtype, uint8_t* state, cudaStream_t st) {
    k_contain_init<<<cdiv(rd.r_hi - rd.r_lo, 256), 256, 0, st>>>(rv, rd, active0, rtype, state);
    g_launches += 1;
}

void launch_contain_step(const RecView& rv, const ReadView& rd, const uint8_t* active0,
                         const uint8_t* rtype, uint8_t* state, int* remaining, cudaStream_t st) {
    cudaMemsetAsync(remaining, 0, sizeof(int), st);
    k_contain_step<<<cdiv(rd.r_hi - rd.r_lo, 256), 256, 0, st>>>(rv, rd, active0, rtype, state,
                                                                 remaining);
    g_launches += 1;
}

void launch_sort_candidates(const SelectIO& io, cudaStream_t st) {
    k_sort_candidates<<<cdiv(io.n_read, 128), 128, 0, st>>>(io);
    g_launches += 1;
}

void launch_hinge_graph(const RecView& rv, const hg_layout_params& P, const SelectIO& io,
                        cudaStream_t st) {
    cudaMemsetAsync(io.counters, 0, sizeof(int) * 8, st);
    k_hinge_graph<<<cdiv(io.n_read, 128), 128, 0, st>>>(rv, P, io);
    g_launches += 1;
}

void launch_best_extension(const hg_layout_params& P, const SelectIO& io, cudaStream_t st) {
    cudaMemsetAsync(io.counters + 2, 0, sizeof(int), st);
    k_best_extension<<<cdiv(io.n_read, 128), 128, 0, st>>>(P, io);
    g_launches += 1;
}

}  // namespace hg
