// See adjacency_list.hpp in this directory: stand-in for Boost.Graph's
// connected_components (labels components 0..k-1 in order of first vertex).
#ifndef HG_ORACLE_BOOST_SHIM_CONNECTED_COMPONENTS_HPP
#define HG_ORACLE_BOOST_SHIM_CONNECTED_COMPONENTS_HPP
#include <vector>
#include "adjacency_list.hpp"

namespace boost {

template <class A, class B, class C>
inline int connected_components(const adjacency_list<A, B, C>& g, int* comp) {
    const std::size_t n = g.nbr.size();
    for (std::size_t i = 0; i < n; ++i) comp[i] = -1;
    int label = 0;
    std::vector<std::size_t> stack;
    for (std::size_t s = 0; s < n; ++s) {
        if (comp[s] != -1) continue;
        comp[s] = label;
        stack.push_back(s);
        while (!stack.empty()) {
            std::size_t u = stack.back();
            stack.pop_back();
            for (std::size_t k = 0; k < g.nbr[u].size(); ++k) {
                std::size_t v = g.nbr[u][k];
                if (comp[v] == -1) { comp[v] = label; stack.push_back(v); }
            }
        }
        ++label;
    }
    return label;
}

}  // namespace boost
#endif
