// Minimal stand-in for the two Boost.Graph headers that the reference's
// src/layout/hinging.cpp includes (hinging.cpp:25-26,40,1338,1420,1644-1645).
// Boost is an external, un-vendored dependency of the reference and is not
// installed in this image.  The reference only uses
//     adjacency_list<vecS, vecS, undirectedS> g(n); add_edge(u, v, g);
//     num_vertices(g); connected_components(g, &comp[0]);
// and only the SIZE of each component influences its output, so any correct
// component labelling yields identical results.  This file is test
// infrastructure for building oracle/_ref; it is not part of the product.
#ifndef HG_ORACLE_BOOST_SHIM_ADJACENCY_LIST_HPP
#define HG_ORACLE_BOOST_SHIM_ADJACENCY_LIST_HPP
#include <cstddef>
#include <vector>

namespace boost {

struct vecS {};
struct undirectedS {};

template <class OutEdgeList, class VertexList, class Directed>
class adjacency_list {
public:
    explicit adjacency_list(std::size_t n = 0) : nbr(n) {}
    std::vector<std::vector<std::size_t> > nbr;
};

template <class A, class B, class C>
inline void add_edge(std::size_t u, std::size_t v, adjacency_list<A, B, C>& g) {
    std::size_t need = (u > v ? u : v) + 1;
    if (g.nbr.size() < need) g.nbr.resize(need);
    g.nbr[u].push_back(v);
    g.nbr[v].push_back(u);
}

template <class A, class B, class C>
inline std::size_t num_vertices(const adjacency_list<A, B, C>& g) {
    return g.nbr.size();
}

}  // namespace boost
#endif
