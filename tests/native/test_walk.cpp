// Differential test of the trace walk of maximal / layout: the product's classify_record
// (hinge_b200/csrc/hg_layout.cu -- its SOURCE TEXT, cut out by tests/test_trace_walk.py into
// classify_record.inc and compiled here for the host behind a few shims) against a plain restatement of
// the reference's ProcessAlignment = trim_overlap + AddTypesAsymmetric
// (/root/reference/src/maximal/maximal.cpp:65-134, /root/reference/src/lib/LAInterface.cpp:4552-4683,
// 4721-4781) on random matches, masks and traces.  The product walks the trace as a counting loop with
// early exits; what must agree: active, type, and for active matches the trimmed coordinates and weight.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <random>
#include <vector>

using std::max;
using std::min;
#define __device__
#define __forceinline__ inline
#define __restrict__
template <class T>
static inline T __ldg(const T* p) { return *p; }
struct int2 { int x, y; };
#include "hg_params.h"
struct RecView {
    int64_t novl;
    const int32_t *aread, *bread, *abpos, *aepos, *bbpos, *bepos, *flags;
    const int64_t* trace_off;
    const uint8_t* trace;
    int32_t tbytes;
    const int64_t* read_off;
};
#include "classify_record.inc"

namespace ref {
struct Ov {
    int as, ae, bs, be, comp;
    bool active = true;
    int ras = 0, rae = 0, rbs = 0, rbe = 0, eas = 0, eae = 0, ebs = 0, ebe = 0;
    int type = HG_UNDEFINED, weight = 0;
};
typedef std::pair<int, int> PII;
// LAInterface.cpp:4552-4683
static void trim_overlap(Ov* o, const std::vector<int>& tr) {
    o->ebs = o->bs; o->ebe = o->be; o->eas = o->as; o->eae = o->ae;
    std::vector<PII> tp;
    tp.push_back(PII(o->as, o->comp ? o->be : o->bs));
    const int sign = 1 - 2 * o->comp;
    int cur = o->as;
    const int tlen = (int)tr.size();
    for (int j = 0; j < tlen / 2 - 1; j++) {
        if (cur % 100 != 0) cur = int(ceil(cur / 100.0)) * 100; else cur += 100;
        tp.push_back(PII(cur, tp.back().second + sign * tr[2 * j + 1]));
    }
    tp.push_back(PII(o->ae, o->comp ? o->bs : o->be));
    int start_idx = (int)tp.size(), end_idx = 0;
    if (!o->comp) {
        for (int i = 0; i < (int)tp.size(); i++)
            if (tp[i].first >= o->ras && tp[i].second >= o->rbs) { o->eas = tp[i].first; o->ebs = tp[i].second; start_idx = i; break; }
        for (int i = (int)tp.size() - 1; i >= 0; i--)
            if (tp[i].first <= o->rae && tp[i].second <= o->rbe) { o->eae = tp[i].first; o->ebe = tp[i].second; end_idx = i; break; }
    } else {
        for (int i = 0; i < (int)tp.size(); i++)
            if (tp[i].first >= o->ras && tp[i].second <= o->rbe) { o->eas = tp[i].first; o->ebe = tp[i].second; start_idx = i; break; }
        for (int i = (int)tp.size() - 1; i >= 0; i--)
            if (tp[i].first <= o->rae && tp[i].second >= o->rbs) { o->eae = tp[i].first; o->ebs = tp[i].second; end_idx = i; break; }
    }
    if (start_idx >= end_idx) o->active = false;
}
// LAInterface.cpp:4721-4781 (the dangling else of the last branch included)
static void add_types(Ov* o, int max_oh, int min_oh) {
    int al = o->eas - o->ras, ar = o->rae - o->eae, bl = o->ebs - o->rbs, br = o->rbe - o->ebe;
    if (o->comp) std::swap(bl, br);
    if (std::max(al, ar) < max_oh && std::min(bl, br) > min_oh) o->type = HG_BCOVERA;
    else if (std::max(bl, br) < max_oh && std::min(al, ar) > min_oh) o->type = HG_ACOVERB;
    else if (std::min(al, ar) > max_oh) o->type = HG_INTERNAL;
    else if (al <= max_oh) {
        if (br <= max_oh && bl >= max_oh) o->type = HG_BACKWARD;
        else if (br >= max_oh && bl >= max_oh) o->type = HG_BACKWARD_INTERNAL;
    } else if (ar <= max_oh) {
        if (bl <= max_oh && br >= max_oh) o->type = HG_FORWARD;
        else if (bl >= max_oh && br >= max_oh) o->type = HG_FORWARD_INTERNAL;
        else o->type = HG_UNDEFINED;
    }
}
// maximal.cpp:65-134
static void process(Ov* o, const std::vector<int>& tr, PII ma, PII mb, int aln, int theta, int theta2) {
    o->ras = ma.first; o->rae = ma.second; o->rbs = mb.first; o->rbe = mb.second;
    trim_overlap(o, tr);
    if ((o->ebe - o->ebs) < aln || (o->eae - o->eas) < aln || !o->active) {
        o->active = false;
        o->type = HG_NOT_ACTIVE;
    } else {
        add_types(o, theta, theta2);
    }
    o->weight = o->eae - o->eas + o->ebe - o->ebs;
}
}  // namespace ref

int main(int argc, char** argv) {
    const long cases = argc > 1 ? atol(argv[1]) : 400000;
    std::mt19937_64 rng(20261017);
    auto U = [&](int lo, int hi) { return lo + (int)(rng() % (uint64_t)(hi - lo + 1)); };
    long active = 0, bad = 0, types[16] = {0};
    for (long c = 0; c < cases; c++) {
        const int wide = c % 7 == 0;                 // tspace > 125: 16-bit trace values
        const int unit = wide ? 2 : 1;
        const int as = U(0, 4000), span = U(150, 7000), ae = as + span, comp = U(0, 1);
        int inner = (ae - 1) / 100 - as / 100;       // trace points strictly inside, as daligner writes them
        if (c % 11 == 0) inner = max(0, inner + U(-2, 2));   // and not quite consistent ones
        const int bs = U(0, 4000);
        std::vector<int> tr(2 * (inner + 1));
        int bsum = 0;
        for (int j = 0; j <= inner; j++) {
            tr[2 * j] = U(0, 30);
            tr[2 * j + 1] = c % 5 == 0 ? U(0, wide ? 600 : 255) : U(80, 120);
            if (j < inner) bsum += tr[2 * j + 1];
        }
        const int be = bs + max(1, bsum + tr[2 * inner + 1] + (c % 13 == 0 ? U(-150, 150) : 0));
        const int alen = ae + U(0, 3000), blen = be + U(0, 3000);
        // masks: anything from empty over tight to the whole read
        auto mask_of = [&](int len, int s, int e) {
            switch (U(0, 5)) {
                case 0: return ref::PII(0, len);
                case 1: return ref::PII(0, 0);
                case 2: return ref::PII(U(0, len), U(0, len));
                case 3: return ref::PII(max(0, s - U(0, 400)), min(len, e + U(0, 400)));
                case 4: return ref::PII(min(len, s + U(0, 600)), max(0, e - U(0, 600)));
                default: return ref::PII(40 * U(0, len / 40), 40 * U(0, len / 40));
            }
        };
        const ref::PII ma = mask_of(alen, as, ae), mb = mask_of(blen, bs, be);
        hg_layout_params P;
        memset(&P, 0, sizeof P);
        P.aln_threshold = c % 3 == 0 ? 0 : U(0, 2500);
        P.theta = U(0, 600);
        P.theta2 = U(0, 300);

        ref::Ov o;
        o.as = as; o.ae = ae; o.bs = bs; o.be = be; o.comp = comp;
        ref::process(&o, tr, ma, mb, P.aln_threshold, P.theta, P.theta2);

        // the product's view: struct of arrays, B coordinates as in the file (complemented strand when comp)
        const int32_t aread = 0, bread = 1, abpos = as, aepos = ae, flags = comp;
        const int32_t bbpos = comp ? blen - be : bs, bepos = comp ? blen - bs : be;
        std::vector<uint8_t> raw(tr.size() * unit + 8, 0);
        for (size_t i = 0; i < tr.size(); i++) {
            if (wide) { const uint16_t v = (uint16_t)tr[i]; memcpy(&raw[2 * i], &v, 2); }
            else raw[i] = (uint8_t)tr[i];
        }
        const int64_t toff[2] = {0, (int64_t)tr.size() * unit};
        const int rlen[2] = {alen, blen};
        const int2 mask[2] = {{ma.first, ma.second}, {mb.first, mb.second}};
        RecView rv;
        memset(&rv, 0, sizeof rv);
        rv.novl = 1; rv.aread = &aread; rv.bread = &bread; rv.abpos = &abpos; rv.aepos = &aepos;
        rv.bbpos = &bbpos; rv.bepos = &bepos; rv.flags = &flags; rv.trace_off = toff; rv.trace = raw.data();
        rv.tbytes = unit;
        const Match m = classify_record(rv, rlen, mask, 0, 0, 1, P);

        bool same = m.active == o.active && m.type == o.type;
        if (same && o.active)
            same = m.eas == o.eas && m.eae == o.eae && m.ebs == o.ebs && m.ebe == o.ebe && m.weight == o.weight;
        if (!same && bad++ < 5)
            fprintf(stderr, "case %ld: as %d ae %d bs %d be %d comp %d inner %d maskA (%d,%d) maskB (%d,%d): "
                    "product active %d type %d (%d,%d,%d,%d) reference active %d type %d (%d,%d,%d,%d)\n",
                    c, as, ae, bs, be, comp, inner, ma.first, ma.second, mb.first, mb.second, (int)m.active, m.type,
                    m.eas, m.eae, m.ebs, m.ebe, (int)o.active, o.type, o.eas, o.eae, o.ebs, o.ebe);
        active += o.active;
        types[o.type & 15]++;
    }
    printf("cases %ld active %ld mismatches %ld | forward %ld backward %ld acoverb %ld bcovera %ld undefined %ld "
           "internal %ld not_active %ld fwd_internal %ld bwd_internal %ld\n", cases, active, bad, types[0], types[1],
           types[2], types[3], types[4], types[5], types[6], types[12], types[13]);
    return bad ? 1 : (active < cases / 50 ? 2 : 0);
}
