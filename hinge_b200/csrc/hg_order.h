// Order-exact restatements of two libstdc++ routines whose implementation-
// defined element order leaks into the reference's results:
//
//   * std::sort (introsort; unstable).  The reference sorts every pile-up with
//     compare_overlap (filter.cpp:565-567), the hinge-call end lists with
//     pairAscend/pairDescend (filter.cpp:914,1010), each (A,B) pair's overlaps
//     (maximal.cpp:647-654,791; hinging.cpp:534) and the extension candidates
//     with compare_overlap_weight (hinging.cpp:1066-1071).  Ties are common
//     (64 % of reads), so a kernel that wants the reference's
//     bytes has to reproduce the exact permutation, not just "a" sorted order.
//   * std::unordered_map<int, T> iteration order (maximal.cpp:789,
//     hinging.cpp:532): decides the pre-sort order of candidate extensions.
//
// Both are restated from the published algorithm of GCC 13's libstdc++
// (bits/stl_algo.h: __introsort_loop / __final_insertion_sort / __heap_select;
// bits/hashtable.h + hashtable_policy.h: _Prime_rehash_policy), compile for
// host and device, and are checked element-for-element against the real
// library in tests/test_order_exact.py.
#ifndef HG_ORDER_H
#define HG_ORDER_H
#include <stdint.h>

#if defined(__CUDACC__)
#define HG_HD __host__ __device__ __forceinline__
#else
#define HG_HD inline
#endif

namespace hg {

// ---------------------------------------------------------------- std::sort

template <class T, class Less>
HG_HD void os_unguarded_linear_insert(T* a, int last, Less less) {
    T val = a[last];
    int next = last - 1;
    while (less(val, a[next])) {
        a[last] = a[next];
        last = next;
        --next;
    }
    a[last] = val;
}

template <class T, class Less>
HG_HD void os_insertion_sort(T* a, int first, int last, Less less) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (less(a[i], a[first])) {
            T val = a[i];
            for (int k = i; k > first; --k) a[k] = a[k - 1];
            a[first] = val;
        } else {
            os_unguarded_linear_insert(a, i, less);
        }
    }
}

template <class T, class Less>
HG_HD void os_push_heap(T* a, int first, int hole, int top, T value, Less less) {
    int parent = (hole - 1) / 2;
    while (hole > top && less(a[first + parent], value)) {
        a[first + hole] = a[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    a[first + hole] = value;
}

template <class T, class Less>
HG_HD void os_adjust_heap(T* a, int first, int hole, int len, T value, Less less) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (less(a[first + child], a[first + (child - 1)])) child--;
        a[first + hole] = a[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        a[first + hole] = a[first + (child - 1)];
        hole = child - 1;
    }
    os_push_heap(a, first, hole, top, value, less);
}

// std::__partial_sort(first, last, last): make_heap + sort_heap
template <class T, class Less>
HG_HD void os_heap_sort(T* a, int first, int last, Less less) {
    const int len = last - first;
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            T value = a[first + parent];
            os_adjust_heap(a, first, parent, len, value, less);
            if (parent == 0) break;
            parent--;
        }
    }
    while (last - first > 1) {
        --last;
        T value = a[last];
        a[last] = a[first];
        os_adjust_heap(a, first, 0, last - first, value, less);
    }
}

template <class T>
HG_HD void os_swap(T* a, int i, int j) {
    T t = a[i];
    a[i] = a[j];
    a[j] = t;
}

// Exactly the permutation std::sort(a, a + n, less) produces (GCC 13).
template <class T, class Less>
HG_HD void std_sort_exact(T* a, int n, Less less) {
    if (n <= 0) return;
    // __introsort_loop; the recursion on the right part is replaced by an
    // explicit stack (disjoint ranges commute, so the result is identical)
    int stack_first[64], stack_last[64], stack_depth[64];
    int sp = 0;
    int lg = 0;
    for (unsigned v = (unsigned)n; v > 1; v >>= 1) lg++;
    stack_first[0] = 0;
    stack_last[0] = n;
    stack_depth[0] = 2 * lg;
    sp = 1;
    while (sp > 0) {
        --sp;
        int first = stack_first[sp], last = stack_last[sp], depth = stack_depth[sp];
        while (last - first > 16) {
            if (depth == 0) {
                os_heap_sort(a, first, last, less);
                break;
            }
            --depth;
            // __move_median_to_first(first, first+1, mid, last-1)
            const int mid = first + (last - first) / 2;
            const int x = first + 1, y = mid, z = last - 1;
            if (less(a[x], a[y])) {
                if (less(a[y], a[z])) os_swap(a, first, y);
                else if (less(a[x], a[z])) os_swap(a, first, z);
                else os_swap(a, first, x);
            } else if (less(a[x], a[z])) {
                os_swap(a, first, x);
            } else if (less(a[y], a[z])) {
                os_swap(a, first, z);
            } else {
                os_swap(a, first, y);
            }
            // __unguarded_partition(first+1, last, pivot = first)
            int lo = first + 1, hi = last;
            while (true) {
                while (less(a[lo], a[first])) ++lo;
                --hi;
                while (less(a[first], a[hi])) --hi;
                if (!(lo < hi)) break;
                os_swap(a, lo, hi);
                ++lo;
            }
            const int cut = lo;
            stack_first[sp] = cut;  // right part: later
            stack_last[sp] = last;
            stack_depth[sp] = depth;
            ++sp;
            last = cut;
        }
    }
    // __final_insertion_sort
    if (n > 16) {
        os_insertion_sort(a, 0, 16, less);
        for (int i = 16; i != n; ++i) os_unguarded_linear_insert(a, i, less);
    } else {
        os_insertion_sort(a, 0, n, less);
    }
}


// ---------------------------------------------------------------- std::unordered_map<int, T> order
//
// Iteration order of a default-constructed std::unordered_map<int, T> after inserting n DISTINCT
// non-negative keys one at a time (operator[]), as GCC 13's libstdc++ does it (bits/hashtable.h):
//   * std::hash<int> is the identity, bucket = key % bucket_count
//   * a new node goes to the FRONT of its bucket's run in the singly linked element list; into an
//     empty bucket it goes to the front of the whole list (_M_insert_bucket_begin)
//   * before the insertion that would exceed the load factor the table is rehashed to the next
//     bucket count of _Prime_rehash_policy; the rehash walks the list and re-inserts every node
//     by the same two rules (_M_rehash_aux, unique keys)
// The bucket-count schedule (1, 13, 29, 59, 127, 257, ...: "next prime" table of the library's
// binary) is not restated: hash_growth_schedule() reads it off the real container at run time
// (grow_at[i] = number of elements whose insertion makes the count grow_bkt[i]).
// next[n] and bucket[max bucket count] are scratch; out[n] receives the indices (insertion
// numbers) in iteration order.
constexpr int kHashBeforeBegin = -2, kHashEmpty = -1;

template <class KeyAt>
HG_HD void hash_iteration_order(KeyAt key_at, int n, const int* grow_at, const int* grow_bkt, int ngrow,
                                int* next, int* bucket, int* out) {
    int head = -1, nb = 1, gi = 0;
    bucket[0] = kHashEmpty;
    for (int k = 0; k < n; k++) {
        while (gi < ngrow && grow_at[gi] <= k + 1) {  // rehash before linking element k + 1
            const int newb = grow_bkt[gi++];
            for (int i = 0; i < newb; i++) bucket[i] = kHashEmpty;
            int p = head, bbegin = 0;
            head = -1;
            while (p >= 0) {
                const int nx = next[p];
                const int b = key_at(p) % newb;
                if (bucket[b] == kHashEmpty) {
                    next[p] = head;
                    head = p;
                    bucket[b] = kHashBeforeBegin;
                    if (next[p] >= 0) bucket[bbegin] = p;
                    bbegin = b;
                } else if (bucket[b] == kHashBeforeBegin) {
                    next[p] = head;
                    head = p;
                } else {
                    next[p] = next[bucket[b]];
                    next[bucket[b]] = p;
                }
                p = nx;
            }
            nb = newb;
        }
        const int b = key_at(k) % nb;
        if (bucket[b] == kHashEmpty) {
            next[k] = head;
            head = k;
            if (next[k] >= 0) bucket[key_at(next[k]) % nb] = k;
            bucket[b] = kHashBeforeBegin;
        } else if (bucket[b] == kHashBeforeBegin) {
            next[k] = head;
            head = k;
        } else {
            next[k] = next[bucket[b]];
            next[bucket[b]] = k;
        }
    }
    int i = 0;
    for (int p = head; p >= 0; p = next[p]) out[i++] = p;
}

#if defined(__CUDACC__)
// ------------------------------------------------- std::sort, one warp, same result
//
// warp_sort_exact leaves a[0..n) in exactly the permutation std::sort would, but
// spreads the work over the 32 lanes of the calling warp (all lanes must call it
// with identical arguments):
//   * Hoare partition: with G = positions (ascending) whose element is not less
//     than the pivot and L = positions (descending) whose element is not greater,
//     the sequential scan swaps G[i] <-> L[i] while G[i] < L[i] and returns
//     min(G[m], L[m-1]) (elements between the two cursors are never touched, so
//     the pairs can be read off the unpartitioned range).  Flags by ballot, pair
//     count by a warp reduction, swaps in parallel.
//   * the final insertion sort is a stable sort in which no element leaves its
//     <= 16-element block (blocks are mutually ordered after the introsort loop):
//     every lane ranks its elements inside a +-15 window.
//   * median-of-3 and the heap-sort fallback (depth limit) stay on lane 0.
// Scratch: idx_g / idx_l hold n indices each (16-bit ones do for n < 65536), tmp holds n elements.
template <class T, class Less, class Idx>
__device__ void warp_sort_exact(T* a, int n, Less less, Idx* idx_g, Idx* idx_l, T* tmp) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    if (n <= 1) return;
    int stack_first[64], stack_last[64], stack_depth[64];
    int sp = 1, lg = 0;
    for (unsigned v = (unsigned)n; v > 1; v >>= 1) lg++;
    stack_first[0] = 0;
    stack_last[0] = n;
    stack_depth[0] = 2 * lg;
    while (sp > 0) {
        --sp;
        int first = stack_first[sp], last = stack_last[sp], depth = stack_depth[sp];
        while (last - first > 16) {
            if (depth == 0) {
                if (lane == 0) os_heap_sort(a, first, last, less);
                __syncwarp();
                break;
            }
            --depth;
            if (lane == 0) {  // __move_median_to_first(first, first+1, mid, last-1)
                const int mid = first + (last - first) / 2;
                const int x = first + 1, y = mid, z = last - 1;
                if (less(a[x], a[y])) {
                    if (less(a[y], a[z])) os_swap(a, first, y);
                    else if (less(a[x], a[z])) os_swap(a, first, z);
                    else os_swap(a, first, x);
                } else if (less(a[x], a[z])) {
                    os_swap(a, first, x);
                } else if (less(a[y], a[z])) {
                    os_swap(a, first, z);
                } else {
                    os_swap(a, first, y);
                }
            }
            __syncwarp();
            const T pivot = a[first];
            int ng = 0, nl = 0;
            for (int base = first + 1; base < last; base += 32) {
                const int p = base + lane;
                bool ge = false, le = false;
                if (p < last) {
                    const T v = a[p];
                    ge = !less(v, pivot);
                    le = !less(pivot, v);
                }
                const unsigned mg = __ballot_sync(0xffffffffu, ge), ml = __ballot_sync(0xffffffffu, le);
                if (ge) idx_g[ng + __popc(mg & lt)] = (Idx)p;
                if (le) idx_l[nl + __popc(ml & lt)] = (Idx)p;
                ng += __popc(mg);
                nl += __popc(ml);
            }
            __syncwarp();
            const int lim = ng < nl ? ng : nl;
            int cnt = 0;
            for (int i = lane; i < lim; i += 32) cnt += (int)idx_g[i] < (int)idx_l[nl - 1 - i] ? 1 : 0;
            const int m = __reduce_add_sync(0xffffffffu, cnt);
            for (int i = lane; i < m; i += 32) os_swap(a, (int)idx_g[i], (int)idx_l[nl - 1 - i]);
            const int prev_hi = m > 0 ? (int)idx_l[nl - m] : last;
            const int cut = (m < ng && (int)idx_g[m] < prev_hi) ? (int)idx_g[m] : prev_hi;
            __syncwarp();
            stack_first[sp] = cut;
            stack_last[sp] = last;
            stack_depth[sp] = depth;
            ++sp;
            last = cut;
        }
    }
    // __final_insertion_sort as a windowed stable rank
    for (int p = lane; p < n; p += 32) {
        const T v = a[p];
        int np = p;
        const int lo = p - 15 > 0 ? p - 15 : 0, hi = p + 15 < n - 1 ? p + 15 : n - 1;
        for (int j = lo; j < p; j++) np -= less(v, a[j]) ? 1 : 0;
        for (int j = p + 1; j <= hi; j++) np += less(a[j], v) ? 1 : 0;
        tmp[np] = v;
    }
    __syncwarp();
    for (int p = lane; p < n; p += 32) a[p] = tmp[p];
    __syncwarp();
}

// ------------------------------------------------- std::sort, one CTA, same result
//
// The introsort loop is a tree: after a partition the two parts are independent (disjoint
// ranges commute, see std_sort_exact), so the warps of a CTA work on different sub-ranges at the
// same time and the critical path shrinks from the ~n/8 partition steps of one warp to the
// depth of the tree.  Ranges wait in a small shared-memory stack guarded by a spin lock that
// only lane 0 of a warp ever touches.  All threads of the CTA call this with identical
// arguments; idx_g / idx_l hold n ints each (a range uses its own [first, last) slice), tmp n
// elements.
constexpr int kCtaSortStack = 160;   // shared stack of pending sub-ranges
constexpr int kCtaSortLocal = 64;    // per-warp overflow stack: one entry per level of the introsort
                                     // recursion at most, and the depth limit is 2 log2(n) <= 62
struct CtaSortState {
    int first[kCtaSortStack], last[kCtaSortStack], depth[kCtaSortStack];
    int top, busy, lock;
};

template <class T, class Less>
__device__ void cta_sort_exact(T* a, int n, Less less, int* idx_g, int* idx_l, T* tmp, CtaSortState* st) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    if (n <= 1) return;  // uniform
    if (threadIdx.x == 0) {
        int lg = 0;
        for (unsigned v = (unsigned)n; v > 1; v >>= 1) lg++;
        st->first[0] = 0;
        st->last[0] = n;
        st->depth[0] = 2 * lg;
        st->top = 1;
        st->busy = 0;
        st->lock = 0;
    }
    __syncthreads();
    auto lock = [&]() {
        while (atomicCAS(&st->lock, 0, 1) != 0) {
        }
        __threadfence_block();
    };
    auto unlock = [&]() {
        __threadfence_block();
        atomicExch(&st->lock, 0);
    };
    for (;;) {
        int first = 0, last = 0, depth = 0, state = 0;  // state: 1 = got a range, 2 = all done
        if (lane == 0) {
            lock();
            if (st->top > 0) {
                const int s = --st->top;
                first = st->first[s];
                last = st->last[s];
                depth = st->depth[s];
                st->busy++;
                state = 1;
            } else if (st->busy == 0) {
                state = 2;
            }
            unlock();
        }
        state = __shfl_sync(0xffffffffu, state, 0);
        if (state == 2) break;
        if (state == 0) {
            __nanosleep(200);
            continue;
        }
        first = __shfl_sync(0xffffffffu, first, 0);
        last = __shfl_sync(0xffffffffu, last, 0);
        depth = __shfl_sync(0xffffffffu, depth, 0);
        // the range was written by another warp before it pushed it: lane 0's lock acquire + fence
        // orders those writes before lane 0, this orders them before the other 31 lanes
        __syncwarp();
        // right-hand parts that found the shared stack full wait here and are done by this warp
        int lf[kCtaSortLocal], ll[kCtaSortLocal], ld[kCtaSortLocal], lsp = 0;
      next_range:
        while (last - first > 16) {
            if (depth == 0) {
                if (lane == 0) os_heap_sort(a, first, last, less);
                __syncwarp();
                break;
            }
            --depth;
            if (lane == 0) {  // __move_median_to_first(first, first+1, mid, last-1)
                const int mid = first + (last - first) / 2;
                const int x = first + 1, y = mid, z = last - 1;
                if (less(a[x], a[y])) {
                    if (less(a[y], a[z])) os_swap(a, first, y);
                    else if (less(a[x], a[z])) os_swap(a, first, z);
                    else os_swap(a, first, x);
                } else if (less(a[x], a[z])) {
                    os_swap(a, first, x);
                } else if (less(a[y], a[z])) {
                    os_swap(a, first, z);
                } else {
                    os_swap(a, first, y);
                }
            }
            __syncwarp();
            const T pivot = a[first];
            int* const pg = idx_g + first;  // at most last - first - 1 entries each
            int* const pl = idx_l + first;
            int ng = 0, nl = 0;
            for (int base = first + 1; base < last; base += 32) {
                const int p = base + lane;
                bool ge = false, le = false;
                if (p < last) {
                    const T v = a[p];
                    ge = !less(v, pivot);
                    le = !less(pivot, v);
                }
                const unsigned mg = __ballot_sync(0xffffffffu, ge), ml = __ballot_sync(0xffffffffu, le);
                if (ge) pg[ng + __popc(mg & lt)] = p;
                if (le) pl[nl + __popc(ml & lt)] = p;
                ng += __popc(mg);
                nl += __popc(ml);
            }
            __syncwarp();
            const int lim = ng < nl ? ng : nl;
            int cnt = 0;
            for (int i = lane; i < lim; i += 32) cnt += pg[i] < pl[nl - 1 - i] ? 1 : 0;
            const int m = __reduce_add_sync(0xffffffffu, cnt);
            for (int i = lane; i < m; i += 32) os_swap(a, pg[i], pl[nl - 1 - i]);
            const int prev_hi = m > 0 ? pl[nl - m] : last;
            const int cut = (m < ng && pg[m] < prev_hi) ? pg[m] : prev_hi;
            __syncwarp();
            if (last - cut > 16) {  // right part: for whoever is free (parts of <= 16 need nothing here)
                int pushed = 0;
                if (lane == 0) {
                    lock();
                    if (st->top < kCtaSortStack) {
                        const int s = st->top++;
                        st->first[s] = cut;
                        st->last[s] = last;
                        st->depth[s] = depth;
                        pushed = 1;
                    }
                    unlock();
                }
                pushed = __shfl_sync(0xffffffffu, pushed, 0);
                if (!pushed) {  // shared stack full: keep it (lsp < kCtaSortLocal, see above)
                    lf[lsp] = cut;
                    ll[lsp] = last;
                    ld[lsp] = depth;
                    lsp++;
                }
            }
            last = cut;
        }
        if (lsp > 0) {
            lsp--;
            first = lf[lsp];
            last = ll[lsp];
            depth = ld[lsp];
            goto next_range;
        }
        if (lane == 0) {
            lock();
            st->busy--;
            unlock();
        }
    }
    __syncthreads();
    // __final_insertion_sort as a windowed stable rank (see warp_sort_exact), all threads
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
        const T v = a[p];
        int np = p;
        const int lo = p - 15 > 0 ? p - 15 : 0, hi = p + 15 < n - 1 ? p + 15 : n - 1;
        for (int j = lo; j < p; j++) np -= less(v, a[j]) ? 1 : 0;
        for (int j = p + 1; j <= hi; j++) np += less(a[j], v) ? 1 : 0;
        tmp[np] = v;
    }
    __syncthreads();
    for (int p = threadIdx.x; p < n; p += blockDim.x) a[p] = tmp[p];
    __syncthreads();
}
#endif

}  // namespace hg
#endif
