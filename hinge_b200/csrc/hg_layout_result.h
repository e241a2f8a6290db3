// Host-side result of hg_layout, kept in the context for the file writers.
#ifndef HG_LAYOUT_RESULT_H
#define HG_LAYOUT_RESULT_H
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/hinge_b200.h"
#include "hg_layout.h"

namespace hg {

struct LayoutResult {
    int n_read = 0;
    std::vector<int32_t> mask;      // 2 per read
    std::vector<uint8_t> active;    // reads that take part in the layout
    std::vector<int> garbage;       // .garbage.txt
    std::vector<Cand> cands;        // sorted by (a, b, rank)
    std::vector<int4> ranges;       // per read: forward [x,y) and backward [z,w) slices of order
    std::vector<int> order;         // candidate indices in weight order
    std::vector<int64_t> hin_off, kil_off;
    std::vector<int> hin_pos, hin_type, kil_pos, kil_type;
    std::vector<uint8_t> hin_alive;
    std::vector<GraphRec> graph;    // .hgraph, in output order
    std::vector<SkipRec> skips;     // .edges.skipped, in output order
    std::vector<int2> chosen;       // per read x 2: (candidate, hinge_pos)
    // the chosen candidates themselves (always fetched): slot -> (2 * read + direction, hinge_pos)
    std::vector<Cand> edge_cands;
    std::vector<int2> edge_ref;
    int n_contained = 0;            // "[contained] Should not happen" events (hinging.cpp:598-601)
    // the full candidate lists stay on the device until a file writer asks for them (materialize)
    Cand* d_cands = nullptr;
    int* d_order = nullptr;
    int4* d_ranges = nullptr;
    int64_t n_cand_slots = 0;
    bool materialized = false;

    void fill_edge(int cand, int hinge_pos, hg_edge* e) const { fill_edge(cands[cand], hinge_pos, e); }
    void fill_edge(const Cand& c, int hinge_pos, hg_edge* e) const {
        e->a = c.a; e->b = c.b; e->length = c.length; e->comp = c.comp; e->type = c.type;
        e->weight = c.weight;
        e->eff_a[0] = c.eas; e->eff_a[1] = c.eae; e->eff_b[0] = c.ebs; e->eff_b[1] = c.ebe;
        e->read_a[0] = mask[2 * (size_t)c.a]; e->read_a[1] = mask[2 * (size_t)c.a + 1];
        e->read_b[0] = mask[2 * (size_t)c.b]; e->read_b[1] = mask[2 * (size_t)c.b + 1];
        e->raw_a[0] = c.as; e->raw_a[1] = c.ae; e->raw_b[0] = c.bs; e->raw_b[1] = c.be;
        e->hinge_pos = hinge_pos;
    }
};

// The result of the last hg_layout with the candidate lists copied to the host (first call).
const LayoutResult* layout_result(hg_ctx* c);

}  // namespace hg
#endif
