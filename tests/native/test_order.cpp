// Checks hg::std_sort_exact against the real std::sort element for element on
// tie-heavy inputs, including median-of-3 killer sequences that drive
// introsort into its heap-sort fallback.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "../../hinge_b200/csrc/hg_order.h"

struct E {
    int key, id;
};
struct LessKey {
    bool operator()(const E& a, const E& b) const { return a.key < b.key; }
};
struct GreaterKey {
    bool operator()(const E& a, const E& b) const { return a.key > b.key; }
};

static uint64_t s = 88172645463325252ull;
static uint32_t rnd() {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    return (uint32_t)(s >> 11);
}

template <class Less>
static int check(std::vector<E> v, Less less, const char* what) {
    std::vector<E> a = v, b = v;
    std::sort(a.begin(), a.end(), less);
    hg::std_sort_exact(b.data(), (int)b.size(), less);
    for (size_t i = 0; i < a.size(); i++)
        if (a[i].key != b[i].key || a[i].id != b[i].id) {
            printf("MISMATCH %s n=%zu at %zu\n", what, v.size(), i);
            return 1;
        }
    return 0;
}

// Musser's median-of-3 killer for an ascending introsort
static std::vector<E> killer(int n) {
    std::vector<E> v(n);
    int k = n / 2;
    for (int i = 0; i < k; i++) {
        v[i].key = (i % 2 == 0) ? i + 1 : k + i + (k % 2 == 0 ? 0 : 1);
        v[k + i].key = 2 * (i + 1);
    }
    if (n % 2) v[n - 1].key = n;
    for (int i = 0; i < n; i++) v[i].id = i;
    return v;
}

int main() {
    int bad = 0, cases = 0;
    const int sizes[] = {0, 1, 2, 3, 15, 16, 17, 18, 31, 32, 33, 63, 64, 65, 100, 127, 200, 257,
                         500, 1000, 2048, 5000, 20000};
    for (int n : sizes)
        for (int distinct : {1, 2, 3, 5, 17, 100, 1000, 1 << 30})
            for (int rep = 0; rep < 6; rep++) {
                std::vector<E> v(n);
                for (int i = 0; i < n; i++) {
                    v[i].key = (int)(rnd() % (uint32_t)distinct);
                    v[i].id = i;
                }
                if (rep == 4) std::sort(v.begin(), v.end(), LessKey());
                if (rep == 5) std::sort(v.begin(), v.end(), GreaterKey());
                bad += check(v, LessKey(), "asc");
                bad += check(v, GreaterKey(), "desc");
                cases += 2;
            }
    for (int n : {64, 100, 1000, 4096, 30000, 100001}) {
        std::vector<E> v = killer(n);
        bad += check(v, LessKey(), "killer-asc");
        for (auto& e : v) e.key = -e.key;
        bad += check(v, GreaterKey(), "killer-desc");
        // killer with ties
        for (auto& e : v) e.key /= 3;
        bad += check(v, GreaterKey(), "killer-ties");
        cases += 3;
    }
    printf("%s: %d cases, %d mismatches\n", bad ? "FAIL" : "OK", cases, bad);
    return bad ? 1 : 0;
}
