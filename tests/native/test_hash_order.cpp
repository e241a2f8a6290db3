// hash_iteration_order (hinge_b200/csrc/hg_order.h) against the real std::unordered_map of this
// toolchain: ascending, random and clustered key sets of many sizes.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <unordered_map>
#include <vector>

#include "../../hinge_b200/csrc/hg_order.h"

static void schedule(int max_n, std::vector<int>* at, std::vector<int>* bkt) {
    std::unordered_map<int, int> m;
    size_t cur = m.bucket_count();
    for (int k = 1; k <= max_n; k++) {
        m[k] = 0;
        if (m.bucket_count() != cur) {
            cur = m.bucket_count();
            at->push_back(k);
            bkt->push_back((int)cur);
        }
    }
}

int main() {
    std::vector<int> at, bkt;
    schedule(70000, &at, &bkt);
    int cases = 0;
    unsigned long long seed = 12345;
    auto rnd = [&]() { seed = seed * 6364136223846793005ull + 1442695040888963407ull; return (unsigned)(seed >> 33); };
    for (int n : {0, 1, 2, 3, 11, 12, 13, 14, 28, 29, 30, 58, 59, 60, 100, 126, 127, 128, 200, 257, 258, 541, 600, 1000,
                  1109, 1110, 5000, 33000, 65000})
        for (int mode = 0; mode < 4; mode++) {
            std::vector<int> keys;
            if (mode == 0)
                for (int i = 0; i < n; i++) keys.push_back(i * 3 + 7);
            else if (mode == 1)
                for (int i = 0; i < n; i++) keys.push_back((int)(rnd() % 2000000));
            else if (mode == 2)
                for (int i = 0; i < n; i++) keys.push_back(1000 + i);
            else
                for (int i = 0; i < n; i++) keys.push_back((int)(rnd() % 500000) * 13);
            std::sort(keys.begin(), keys.end());
            keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
            if (mode == 3) std::reverse(keys.begin(), keys.end());
            const int m = (int)keys.size();
            std::unordered_map<int, int> um;
            for (int i = 0; i < m; i++) um[keys[i]] = i;
            std::vector<int> want;
            for (auto it = um.begin(); it != um.end(); ++it) want.push_back(it->second);
            std::vector<int> next(m + 1), bucket(200000), got(m + 1);
            hg::hash_iteration_order([&](int i) { return keys[i]; }, m, at.data(), bkt.data(), (int)at.size(),
                                     next.data(), bucket.data(), got.data());
            got.resize(m);
            if (got != want) {
                printf("MISMATCH n=%d mode=%d\n", m, mode);
                return 1;
            }
            cases++;
        }
    printf("hash order: %d cases identical to std::unordered_map\n", cases);
    return 0;
}
