// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See hinge_oracle.h.
//
// Every function cites the reference lines it restates (paths relative to
// /root/reference/src).  Sorting and hashing go through libstdc++ itself
// (std::sort, std::nth_element, std::unordered_map) because their
// implementation-defined element order is part of the reference's behaviour.
#include "hinge_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <atomic>
#include <functional>
#include <thread>
#include <unordered_map>

namespace oracle {

enum { FORWARD = 0, BACKWARD = 1, ACOVERB = 2, BCOVERA = 3, UNDEFINED = 4, INTERNAL = 5,
       NOT_ACTIVE = 6, FORWARD_INTERNAL = 12, BACKWARD_INTERNAL = 13 };  // LAInterface.h:30-32

// One overlap after ingest: lib/LAInterface.cpp:1583-1629 (B flipped to the
// forward strand for complemented matches).
struct Ov {
    int a, b, as, ae, bs, be, comp;
    int64_t rec;  // row in Data (for the trace)
    bool active = true;
    int ras = 0, rae = 0, rbs = 0, rbe = 0;  // eff_read_*_read_{start,end}
    int eas = 0, eae = 0, ebs = 0, ebe = 0;  // eff_read_*_match_{start,end}
    int type = UNDEFINED, weight = 0, length = 0;
};

// Loops whose iterations are independent (one read, or one record, each) may be spread over
// host threads (Params::threads, used by bench.py's verification leg on 100 M-record sets); every
// iteration computes exactly what the sequential loop computes, so the results do not depend on
// the thread count.
static void parallel_for(int64_t lo, int64_t hi, int threads, const std::function<void(int64_t)>& fn) {
    if (threads <= 1 || hi - lo < 2) {
        for (int64_t i = lo; i < hi; i++) fn(i);
        return;
    }
    std::atomic<int64_t> next(lo);
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(4096, (hi - lo) / (threads * 8)));
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
        th.emplace_back([&]() {
            for (;;) {
                const int64_t b = next.fetch_add(chunk);
                if (b >= hi) break;
                const int64_t e = std::min(hi, b + chunk);
                for (int64_t i = b; i < e; i++) fn(i);
            }
        });
    for (auto& t : th) t.join();
}

static void ingest(const Data& d, std::vector<Ov>* ovs, int threads = 1) {
    ovs->resize((size_t)d.novl);
    parallel_for(0, d.novl, threads, [&](int64_t k) {
        Ov& o = (*ovs)[k];
        o.a = d.aread[k];
        o.b = d.bread[k];
        o.as = d.abpos[k];
        o.ae = d.aepos[k];
        o.comp = d.flags[k] & 1;  // align.h:155 COMP()
        if (!o.comp) {
            o.bs = d.bbpos[k];
            o.be = d.bepos[k];
        } else {
            int blen = d.rlen[o.b];
            o.bs = blen - d.bepos[k];
            o.be = blen - d.bbpos[k];
        }
        o.rec = k;
    });
}

static inline int span(const Ov* o) { return (o->ae - o->as) + (o->be - o->bs); }
// lib/LAInterface.cpp:4884-4889
static bool compare_overlap(const Ov* x, const Ov* y) { return span(x) > span(y); }
// lib/LAInterface.cpp:4921-4923
static bool compare_overlap_weight(const Ov* x, const Ov* y) { return x->weight > y->weight; }
static bool pair_ascend(const PII& x, const PII& y) { return x.first < y.first; }   // :4875
static bool pair_descend(const PII& x, const PII& y) { return x.first > y.first; }  // :4879
static bool compare_event(PII x, PII y) { return x.first < y.first; }               // :4293

// lib/LAInterface.cpp:4298-4320
static void profile_coverage(const std::vector<Ov*>& pile, std::vector<PII>* cov, int reso,
                             int cutoff) {
    std::vector<PII> ev;
    for (size_t i = 0; i < pile.size(); i++) {
        ev.push_back(PII(pile[i]->as + cutoff, 1));
        ev.push_back(PII(pile[i]->ae - cutoff, -1));
    }
    std::sort(ev.begin(), ev.end(), compare_event);
    size_t pos = 0;
    int i = 0, count = 0;
    while (pos < ev.size()) {
        while (pos < ev.size() && ev[pos].first < i * reso) {
            count += ev[pos].second;
            pos++;
        }
        cov->push_back(PII(i * reso, count));
        i++;
    }
}

static inline int trace_at(const Data& d, int64_t rec, int k) {
    int64_t off = d.trace_off[rec];
    if (d.tbytes == 1) return d.trace[off + k];
    uint16_t v;
    memcpy(&v, &d.trace[off + 2 * (int64_t)k], 2);
    return v;
}
static inline int trace_len(const Data& d, int64_t rec) {
    return (int)((d.trace_off[rec + 1] - d.trace_off[rec]) / d.tbytes);
}

// lib/LAInterface.cpp:4552-4683
static void trim_overlap(const Data& d, Ov* o) {
    o->ebs = o->bs; o->ebe = o->be; o->eas = o->as; o->eae = o->ae;
    std::vector<PII> tp;
    tp.push_back(PII(o->as, o->comp ? o->be : o->bs));
    int sign = 1 - 2 * o->comp;
    int cur = o->as;
    int tlen = trace_len(d, o->rec);
    for (int j = 0; j < tlen / 2 - 1; j++) {
        if (cur % 100 != 0)
            cur = int(ceil(cur / 100.0)) * 100;
        else
            cur += 100;
        tp.push_back(PII(cur, tp.back().second + sign * trace_at(d, o->rec, 2 * j + 1)));
    }
    tp.push_back(PII(o->ae, o->comp ? o->bs : o->be));
    int start_idx = (int)tp.size(), end_idx = 0;
    if (!o->comp) {
        for (int i = 0; i < (int)tp.size(); i++)
            if (tp[i].first >= o->ras && tp[i].second >= o->rbs) {
                o->eas = tp[i].first; o->ebs = tp[i].second; start_idx = i;
                break;
            }
        for (int i = (int)tp.size() - 1; i >= 0; i--)
            if (tp[i].first <= o->rae && tp[i].second <= o->rbe) {
                o->eae = tp[i].first; o->ebe = tp[i].second; end_idx = i;
                break;
            }
    } else {
        for (int i = 0; i < (int)tp.size(); i++)
            if (tp[i].first >= o->ras && tp[i].second <= o->rbe) {
                o->eas = tp[i].first; o->ebe = tp[i].second; start_idx = i;
                break;
            }
        for (int i = (int)tp.size() - 1; i >= 0; i--)
            if (tp[i].first <= o->rae && tp[i].second >= o->rbs) {
                o->eae = tp[i].first; o->ebs = tp[i].second; end_idx = i;
                break;
            }
    }
    if (start_idx >= end_idx) o->active = false;
}

// lib/LAInterface.cpp:4721-4781
static void add_types_asymmetric(Ov* o, int max_oh, int min_oh) {
    int al = o->eas - o->ras, ar = o->rae - o->eae;
    int bl = o->ebs - o->rbs, br = o->rbe - o->ebe;
    if (o->comp) std::swap(bl, br);
    if (std::max(al, ar) < max_oh && std::min(bl, br) > min_oh)
        o->type = BCOVERA;
    else if (std::max(bl, br) < max_oh && std::min(al, ar) > min_oh)
        o->type = ACOVERB;
    else if (std::min(al, ar) > max_oh)
        o->type = INTERNAL;
    else if (al <= max_oh) {
        if (br <= max_oh && bl >= max_oh)
            o->type = BACKWARD;
        else if (br >= max_oh && bl >= max_oh)
            o->type = BACKWARD_INTERNAL;
    } else if (ar <= max_oh) {
        if (bl <= max_oh && br >= max_oh)
            o->type = FORWARD;
        else if (bl >= max_oh && br >= max_oh)
            o->type = FORWARD_INTERNAL;
        else
            o->type = UNDEFINED;
    }
}

// maximal/maximal.cpp:65-134 == layout/hinging.cpp:78-147
static bool process_alignment(const Data& d, Ov* o, const std::vector<PII>& mask, const Params& p) {
    bool contained = false;
    o->ras = mask[o->a].first; o->rae = mask[o->a].second;
    o->rbs = mask[o->b].first; o->rbe = mask[o->b].second;
    trim_overlap(d, o);
    if ((o->ebe - o->ebs) < p.aln_threshold || (o->eae - o->eas) < p.aln_threshold || !o->active) {
        o->active = false;
        o->type = NOT_ACTIVE;
    } else {
        add_types_asymmetric(o, p.theta, p.theta2);
        if (o->type == BCOVERA) contained = true;
    }
    o->weight = o->eae - o->eas + o->ebe - o->ebs;
    o->length = o->ae - o->as + o->be - o->bs;
    return contained;
}

// lib/LAInterface.cpp:4498-4546
static int get_matching_position(const Data& d, const Ov* o, int pos_a) {
    if (pos_a < o->as || pos_a > o->ae) return -1;
    int sign = 1 - 2 * o->comp;
    int cur_a = o->as, next_a = cur_a;
    int cur_b = o->comp ? o->be : o->bs;
    int tlen = trace_len(d, o->rec);
    for (int j = 0; j < tlen / 2 - 1; j++) {
        if (cur_a % 100 != 0)
            next_a = int(ceil(cur_a / 100.0)) * 100;
        else
            next_a = cur_a + 100;
        if (next_a >= pos_a) return cur_b + pos_a - cur_a;
        cur_b += sign * trace_at(d, o->rec, 2 * j + 1);
        cur_a = next_a;
    }
    if (cur_a < pos_a) return cur_b + pos_a - cur_a;
    return -2;
}

// ------------------------------------------------------------------ filter

// filter/filter.cpp:340-369
static void qv_masks(const Data& d, std::vector<PII>* qm) {
    qm->assign(d.n_read, PII(0, 0));
    if (!d.has_qv) return;
    for (int i = 0; i < d.n_read; i++) {
        int s = 0, e = 0, max = 0, maxs = 0, maxe = 0;
        int n = (int)(d.qv_off[i + 1] - d.qv_off[i]);
        for (int j = 0; j < n; j++) {
            int good = d.qv[d.qv_off[i] + j] < 40;  // filter.cpp:311
            if (good && j < n - 1) {
                e++;
            } else {
                if (e - s > max) { maxe = e; maxs = s; max = e - s; }
                s = j + 1;
                e = j + 1;
            }
        }
        (*qm)[i] = PII(maxs * d.tspace, maxe * d.tspace);
    }
}

void run_filter(const Data& d, const Params& p, FilterOut* out, FilterCarry* carry) {
    const int n_read = d.n_read;
    const int T = p.threads;
    std::vector<Ov> ovs;
    ingest(d, &ovs, T);
    std::vector<PII> qm;
    qv_masks(d, &qm);
    const bool use_qv = p.use_qv && d.has_qv;  // filter.cpp:409
    int MIN_COV = (carry && carry->started) ? carry->min_cov : p.min_cov;

    out->r_begin = ovs.front().a;  // filter.cpp:516-517
    out->r_end = ovs.back().a;
    const int rb = out->r_begin, re = out->r_end;

    // pile-ups and self alignments, filter.cpp:529-561
    std::vector<std::vector<Ov*>> pile(n_read);
    std::unordered_map<int, std::vector<PII>> self_aln;
    for (size_t k = 0; k < ovs.size(); k++) {
        Ov* o = &ovs[k];
        if (o->a == o->b) {
            o->active = false;
            self_aln[o->a].push_back(PII(o->as, o->ae));
            self_aln[o->a].push_back(PII(o->bs, o->be));
        }
        if (o->active) pile[o->a].push_back(o);
    }
    std::set<int> self_match;
    for (auto& it : self_aln) {
        float cov = 0.0;
        for (size_t i = 0; i < it.second.size(); i++) cov += it.second[i].second - it.second[i].first;
        cov /= float(d.rlen[it.first]);
        if (cov > 4.5 && d.rlen[it.first] > 10000) self_match.insert(it.first);
    }
    parallel_for(0, n_read, T, [&](int64_t i) {  // filter.cpp:565-567
        std::sort(pile[i].begin(), pile[i].end(), compare_overlap);
    });

    // coverage profiles, filter.cpp:588-614
    out->cov0.assign(n_read, std::vector<PII>());
    out->covc.assign(n_read, std::vector<PII>());
    std::vector<std::vector<PII>> cgs(n_read);
    parallel_for(rb, re + 1, T, [&](int64_t i) {
        profile_coverage(pile[i], &out->covc[i], p.reso, p.cut_off);
        profile_coverage(pile[i], &out->cov0[i], p.reso, 0);
        const std::vector<PII>& c = out->cov0[i];
        if (c.size() >= 2)
            for (size_t j = 0; j + 1 < c.size(); j++)
                cgs[i].push_back(PII(c[j].first, c[j + 1].second - c[j].second));
        else
            cgs[i].push_back(PII(0, 0));
    });

    // coverage estimate, filter.cpp:633-678
    {
        std::vector<int> read_cov;
        for (int i = rb; i <= re; i++) {
            if (d.rlen[i] < 5000) continue;
            long rc = 0;
            int slots = 0;
            for (size_t j = 0; j < out->cov0[i].size(); j++) { rc += out->cov0[i][j].second; slots++; }
            read_cov.push_back((int)(rc / std::max(1, slots)));
        }
        size_t mid = read_cov.size() / 2;
        if (mid > 0) std::nth_element(read_cov.begin(), read_cov.begin() + mid, read_cov.end());
        int cov_est = read_cov.empty() ? 0 : read_cov[mid];  // reference: UB when empty
        if (p.est_cov != 0) cov_est = p.est_cov;
        out->cov_est = cov_est;
        if (MIN_COV < cov_est / 3) MIN_COV = cov_est / 3;
        out->min_cov = MIN_COV;
    }

    // masks, filter.cpp:696-789
    out->mask.assign(n_read, PII(0, 0));
    if (carry && carry->started) out->mask = carry->mask;  // reads of earlier parts keep their masks
    out->cmask.assign(n_read, PII(0, 0));
    std::vector<char> cov_flag(n_read, 0), self_flag(n_read, 0);
    parallel_for(rb, re + 1, T, [&](int64_t i) {
        std::vector<PII> cc = out->covc[i];  // thresholded working copy
        for (size_t j = 0; j < cc.size(); j++) {
            cc[j].second -= MIN_COV;
            if (cc[j].second < 0) cc[j].second = 0;
        }
        int start = 0, end = 0, maxlen = 0, maxstart = 0, maxend = 0;
        int sc = 0, ec = 0, msc = 0, mec = 0;
        for (size_t j = 0; j < cc.size(); j++) {
            if (cc[j].second > 0) {
                end = cc[j].first;
                ec = (int)j;
            } else {
                if (end > start && end - start - p.reso > maxlen) {
                    maxlen = end - start - p.reso;
                    maxstart = start + p.reso;
                    maxend = end;
                    msc = sc + 1;
                    mec = ec;
                }
                start = cc[j].first;
                sc = (int)j;
                ec = sc;
                end = start;
            }
        }
        int scov = 0, ecov = 0;
        if (mec - msc + 1 > 20) {
            for (int t = 0; t < 10; t++) {
                scov += cc[msc + t].second + MIN_COV;
                ecov += cc[mec - t].second + MIN_COV;
            }
            scov /= 10;
            ecov /= 10;
        } else {
            int limit = (mec - msc) / 2;
            for (int t = 0; t < limit; t++) {
                scov += cc[msc + t].second + MIN_COV;
                ecov += cc[mec - t].second + MIN_COV;
            }
            if (limit == 0) {
                scov = 0;
                ecov = 0;
            } else {
                scov /= limit;
                ecov /= limit;
            }
        }
        if (p.del_telomere_filter) {
            if (scov >= 10 * ecov || ecov >= 10 * scov) cov_flag[i] = 1;
            if (self_match.count((int)i)) self_flag[i] = 1;
        }
        out->cmask[i] = PII(msc, mec);
        if (use_qv && p.use_coverage)
            out->mask[i] = PII(std::max(maxstart, qm[i].first), std::min(maxend, qm[i].second));
        else if (p.use_coverage && !use_qv)
            out->mask[i] = PII(maxstart, maxend);
        else
            out->mask[i] = qm[i];
    });
    for (int i = rb; i <= re; i++) {
        if (cov_flag[i]) out->cov_flag.push_back(i);
        if (self_flag[i]) out->self_flag.push_back(i);
    }
    const std::vector<PII>& mask = out->mask;

    // repeat annotation + merge, filter.cpp:796-829
    out->repeats.assign(n_read, std::vector<PII>());
    parallel_for(rb, re + 1, T, [&](int64_t i) {
        std::vector<PII>& anno = out->repeats[i];
        for (size_t j = 0; j + 1 < cgs[i].size(); j++) {
            int pos = cgs[i][j].first;
            if (pos >= mask[i].first + p.no_hinge_region && pos <= mask[i].second - p.no_hinge_region) {
                int thr = std::min(std::max((out->cov0[i][j].second + MIN_COV) / p.coverage_fraction,
                                            p.min_rep_thr), p.max_rep_thr);
                if (cgs[i][j].second > thr)
                    anno.push_back(PII(pos, 1));
                else if (cgs[i][j].second < -thr)
                    anno.push_back(PII(pos, -1));
            }
        }
        for (auto it = anno.begin(); it < anno.end();) {
            if (it + 1 < anno.end()) {
                if (it->second == 1 && (it + 1)->second == 1 &&
                    (it + 1)->first - it->first < p.rep_gap)
                    anno.erase(it + 1);
                else if (it->second == -1 && (it + 1)->second == -1 &&
                         (it + 1)->first - it->first < p.rep_gap)
                    it = anno.erase(it);
                else
                    it++;
            } else {
                it++;
            }
        }
    });

    // hinge calls, filter.cpp:838-1070
    out->hinges.assign(n_read, std::vector<PII>());
    const int THETA = p.theta, HBL = p.hinge_bin_length, HTL = p.hinge_tolerance_length;
    parallel_for(rb, re + 1, T, [&](int64_t i) {
        int cs = 0, ns = 0, ne = 0, ce = 0;
        for (size_t j = 0; j < out->cov0[i].size(); j++) {
            int pos = out->cov0[i][j].first;
            if (pos <= mask[i].first + p.no_hinge_region && pos >= mask[i].first) {
                cs += out->cov0[i][j].second;
                ns++;
            }
            if (pos <= mask[i].second && pos >= mask[i].second - p.no_hinge_region) {
                ce += out->cov0[i][j].second;
                ne++;
            }
        }
        float avg_end = (float)ce / ne;
        float avg_start = (float)cs / ns;
        if (std::abs(avg_end - avg_start) < 10) return;

        for (size_t j = 0; j < out->repeats[i].size(); j++) {
            const int apos = out->repeats[i][j].first;
            const bool out_hinge = out->repeats[i][j].second == -1;
            bool bridged = true;
            int support = 0;
            std::vector<PII> ends;
            for (size_t k = 0; k < pile[i].size(); k++) {
                const Ov* o = pile[i][k];
                int lo, ro;
                if (o->comp == 0) {
                    ro = std::max(mask[o->b].second - o->be, 0);
                    lo = std::max(o->bs - mask[o->b].first, 0);
                } else {
                    ro = std::max(o->bs - mask[o->b].first, 0);
                    lo = std::max(mask[o->b].second - o->be, 0);
                }
                if (out_hinge) {
                    if (ro > THETA && o->ae > apos - HTL && o->ae < apos + HTL) {
                        ends.push_back(PII(o->as, lo));
                        support++;
                    }
                } else {
                    if (lo > THETA && o->as > apos - HTL && o->as < apos + HTL) {
                        ends.push_back(PII(o->ae, ro));
                        support++;
                    }
                }
            }
            if (support < p.hinge_min_support) continue;
            if (out_hinge)
                std::sort(ends.begin(), ends.end(), pair_ascend);
            else
                std::sort(ends.begin(), ends.end(), pair_descend);
            int considered = 0, to_end = 0;
            for (int id = 0; id < (int)ends.size(); ++id) {
                int dist_end = out_hinge ? ends[id].first - mask[i].first
                                         : mask[i].second - ends[id].first;
                int dist0 = out_hinge ? ends[id].first - ends[0].first
                                      : ends[0].first - ends[id].first;
                if (dist_end < HBL) {
                    considered++;
                    to_end++;
                    if (to_end > p.hinge_unbridged || (considered > p.hinge_unbridged && dist0 > HBL)) {
                        bridged = false;
                        break;
                    }
                } else if (ends[id].second < THETA) {
                    considered++;
                    if (to_end > p.hinge_unbridged || (considered > p.hinge_unbridged && dist0 > HBL)) {
                        bridged = false;
                        break;
                    }
                } else if (ends[id].second > THETA) {
                    considered++;
                    int id1 = id + 1, pl = 1;
                    while (id1 < (int)ends.size()) {
                        int gap = out_hinge ? ends[id1].first - ends[id].first
                                            : ends[id].first - ends[id1].first;
                        if (gap < HBL) {
                            pl++;
                            id1++;
                        } else {
                            break;
                        }
                    }
                    if (pl > p.hinge_bin_pileup) {
                        bridged = true;
                        break;
                    }
                }
            }
            if (!bridged && support > p.hinge_min_support)
                out->hinges[i].push_back(PII(apos, out_hinge ? -1 : 1));
        }
    });
    if (carry) {
        carry->started = true;
        carry->min_cov = MIN_COV;
        carry->mask = out->mask;
    }
}

// ------------------------------------------------------------------ maximal

typedef std::vector<std::unordered_map<int, std::vector<Ov*>>> PairIndex;

// maximal/maximal.cpp:780-858
void run_maximal(const Data& d, const Params& p, const std::vector<PII>& mask, int r_begin,
                 int r_end, MaximalOut* out) {
    const int n_read = d.n_read;
    std::vector<Ov> ovs;
    ingest(d, &ovs);
    out->active.assign(n_read, 1);
    for (int i = 0; i < n_read; i++)  // maximal.cpp:541-547
        if (mask[i].second - mask[i].first < p.length_threshold) out->active[i] = 0;
    PairIndex idx_ab(n_read);
    for (size_t k = 0; k < ovs.size(); k++) {
        if (ovs[k].a == ovs[k].b) ovs[k].active = false;  // maximal.cpp:616-618
    }
    for (size_t k = 0; k < ovs.size(); k++) idx_ab[ovs[k].a][ovs[k].b] = std::vector<Ov*>();
    for (size_t k = 0; k < ovs.size(); k++) idx_ab[ovs[k].a][ovs[k].b].push_back(&ovs[k]);
    for (int i = 0; i < n_read; i++)  // maximal.cpp:647-654 (first of the two sorts)
        for (auto it = idx_ab[i].begin(); it != idx_ab[i].end(); it++)
            std::sort(it->second.begin(), it->second.end(), compare_overlap);

    for (int i = r_begin; i <= r_end; i++) {
        bool contained = false;
        if (!out->active[i]) continue;
        int containing = 0;
        for (auto it = idx_ab[i].begin(); it != idx_ab[i].end(); it++) {
            std::sort(it->second.begin(), it->second.end(), compare_overlap);
            for (int r = 0; r < 2; r++) {
                if ((int)it->second.size() <= r) break;
                if (r == 1 && !p.use_two_matches) break;
                Ov* o = it->second[r];
                bool ca = process_alignment(d, o, mask, p);
                if (ca) containing = o->b;
                if (out->active[o->b]) contained = contained || ca;
            }
        }
        if (contained) {
            out->active[i] = 0;
            out->contained.push_back(PII(i, containing));
        }
    }
}

// ------------------------------------------------------------------ layout

struct Hinge {
    int pos, type;
    bool active;
};

static Edge make_edge(const Ov* o, int hinge_pos) {
    Edge e;
    e.a = o->a; e.b = o->b; e.length = o->length; e.comp = o->comp; e.type = o->type;
    e.weight = o->weight;
    e.eas = o->eas; e.eae = o->eae; e.ebs = o->ebs; e.ebe = o->ebe;
    e.ras = o->ras; e.rae = o->rae; e.rbs = o->rbs; e.rbe = o->rbe;
    e.as = o->as; e.ae = o->ae; e.bs = o->bs; e.be = o->be;
    e.hinge_pos = hinge_pos;
    return e;
}

void run_layout(const Data& d, const Params& p, const std::vector<PII>& mask,
                const std::vector<char>& maximal, const std::vector<std::vector<PII>>& repeats,
                const std::vector<std::vector<PII>>& hinges_in, LayoutOut* out) {
    const int n_read = d.n_read;
    std::vector<char> active(n_read, 1);
    // hinging.cpp:877-913 telomere kill, :954-960 length filter
    if (p.del_telomeres_layout)
        for (int i = 0; i < n_read; i++)
            if ((int)repeats[i].size() > p.num_events_telomere) active[i] = 0;
    for (int i = 0; i < n_read; i++)
        if (mask[i].second - mask[i].first < p.length_threshold) {
            active[i] = 0;
            out->garbage.push_back(i);
        }
    // GetAlignment, hinging.cpp:398-610
    for (int i = 0; i < n_read; i++) active[i] = active[i] && maximal[i];
    std::vector<Ov> ovs;
    ingest(d, &ovs);
    PairIndex idx_ab(n_read);
    std::vector<std::vector<Ov*>> mf(n_read), mb(n_read);
    const int r_begin = ovs.front().a, r_end = ovs.back().a;
    auto kept = [&](const Ov& o) { return active[o.a] && active[o.b] && p.keep_only_maximal; };
    for (size_t k = 0; k < ovs.size(); k++) {
        if (ovs[k].a == ovs[k].b) ovs[k].active = false;
        if (kept(ovs[k])) idx_ab[ovs[k].a][ovs[k].b] = std::vector<Ov*>();
    }
    for (size_t k = 0; k < ovs.size(); k++)
        if (kept(ovs[k])) idx_ab[ovs[k].a][ovs[k].b].push_back(&ovs[k]);
    for (int i = r_begin; i <= r_end; i++) {
        bool contained = false;
        if (!active[i]) continue;
        for (auto it = idx_ab[i].begin(); it != idx_ab[i].end(); it++) {
            std::sort(it->second.begin(), it->second.end(), compare_overlap);
            for (int r = 0; r < 2; r++) {
                if ((int)it->second.size() <= r) break;
                if (r == 1 && !p.use_two_matches) break;
                Ov* o = it->second[r];
                bool ca = process_alignment(d, o, mask, p);
                if (active[o->b]) contained = contained || ca;
                if (o->type == FORWARD || o->type == FORWARD_INTERNAL)
                    mf[i].push_back(o);
                else if (o->type == BACKWARD || o->type == BACKWARD_INTERNAL)
                    mb[i].push_back(o);
            }
        }
        if (contained) active[i] = 0;  // "[contained] Should not happen", hinging.cpp:598-601
    }
    for (int i = 0; i < n_read; i++)  // hinging.cpp:1066-1071
        if (active[i]) {
            std::sort(mf[i].begin(), mf[i].end(), compare_overlap_weight);
            std::sort(mb[i].begin(), mb[i].end(), compare_overlap_weight);
        }

    // hinge / killed-hinge lists, hinging.cpp:1180-1197
    std::vector<std::vector<Hinge>> hv(n_read), kv(n_read), nk(n_read);
    for (int i = 0; i < n_read; i++) {
        std::set<PII> surviving(hinges_in[i].begin(), hinges_in[i].end());
        for (size_t j = 0; j < hinges_in[i].size(); j++)
            hv[i].push_back(Hinge{hinges_in[i][j].first, hinges_in[i][j].second, true});
        for (size_t j = 0; j < repeats[i].size(); j++)
            if (!surviving.count(repeats[i][j]))
                kv[i].push_back(Hinge{repeats[i][j].first, repeats[i][j].second, false});
    }
    out->killed.assign(n_read, std::vector<PII>());
    for (int i = 0; i < n_read; i++)
        for (size_t j = 0; j < kv[i].size(); j++) out->killed[i].push_back(PII(kv[i][j].type, kv[i][j].pos));

    // kill hinges bridged by matches, hinging.cpp:1262-1321
    for (int i = 0; i < n_read; i++) {
        if (!active[i]) continue;
        for (size_t j = 0; j < mf[i].size(); j++) {
            const Ov* m = mf[i][j];
            if (!m->active || !active[m->b]) continue;
            for (size_t k = 0; k < hv[i].size(); k++)
                if (((m->eas < hv[i][k].pos + p.kill_hinge_internal && m->type == FORWARD_INTERNAL) ||
                     (m->eas < hv[i][k].pos - p.kill_hinge_overlap && m->type == FORWARD)) &&
                    hv[i][k].type == 1)
                    hv[i][k].active = false;
        }
        for (size_t j = 0; j < mb[i].size(); j++) {
            const Ov* m = mb[i][j];
            if (!m->active || !active[m->b]) continue;
            for (size_t k = 0; k < hv[i].size(); k++)
                if (((m->eae > hv[i][k].pos - p.kill_hinge_internal && m->type == BACKWARD_INTERNAL) ||
                     (m->eae > hv[i][k].pos + p.kill_hinge_overlap && m->type == BACKWARD)) &&
                    hv[i][k].type == -1)
                    hv[i][k].active = false;
        }
    }

    // hinge graph, hinging.cpp:1324-1640
    std::vector<int> node_base(n_read + 1, 0);
    for (int i = 0; i < n_read; i++) node_base[i + 1] = node_base[i] + (int)hv[i].size();
    const int num_hinges = node_base[n_read];
    std::vector<int> parent(num_hinges);
    for (int i = 0; i < num_hinges; i++) parent[i] = i;
    auto find = [&](int x) {
        while (parent[x] != x) {
            parent[x] = parent[parent[x]];
            x = parent[x];
        }
        return x;
    };
    auto unite = [&](int x, int y) { parent[find(x)] = find(y); };
    char line[256];
    for (int i = 0; i < n_read; i++) {
        if (!active[i]) continue;
        for (size_t k = 0; k < hv[i].size(); k++) {
            for (int dir = 0; dir < 2; dir++) {
                const std::vector<Ov*>& ms = dir == 0 ? mf[i] : mb[i];
                const int own_type = dir == 0 ? 1 : -1;  // hinge type printed "i first"
                for (size_t j = 0; j < ms.size(); j++) {
                    const Ov* m = ms[j];
                    if (!m->active || !active[m->b]) continue;
                    int pos_b = get_matching_position(d, m, hv[i][k].pos);
                    int req = m->comp ? -hv[i][k].type : hv[i][k].type;
                    int rev = m->comp ? 1 : 0;
                    int b = m->b;
                    for (size_t l = 0; l < hv[b].size(); l++) {
                        if (hv[b][l].pos < pos_b + p.matching_hinge_slack &&
                            hv[b][l].pos > pos_b - p.matching_hinge_slack && req == hv[b][l].type) {
                            unite(node_base[i] + (int)k, node_base[b] + (int)l);
                            if (hv[i][k].type == own_type)
                                snprintf(line, sizeof line, "%d %d %d %d %d %d\n", i, b, hv[i][k].pos,
                                         hv[b][l].pos, 1, rev);
                            else
                                snprintf(line, sizeof line, "%d %d %d %d %d %d\n", b, i, hv[b][l].pos,
                                         hv[i][k].pos, 1, rev);
                            out->hgraph_lines.push_back(line);
                        }
                    }
                    for (size_t l = 0; l < kv[b].size(); l++) {
                        if (kv[b][l].pos < pos_b + p.matching_hinge_slack &&
                            kv[b][l].pos > pos_b - p.matching_hinge_slack) {
                            bool tmatch = req == kv[b][l].type;
                            if (tmatch) {
                                if (hv[i][k].type == own_type)
                                    snprintf(line, sizeof line, "%d %d %d %d %d %d\n", i, b,
                                             hv[i][k].pos, kv[b][l].pos, 0, rev);
                                else
                                    snprintf(line, sizeof line, "%d %d %d %d %d %d\n", b, i,
                                             kv[b][l].pos, hv[i][k].pos, 0, rev);
                                out->hgraph_lines.push_back(line);
                            }
                            // forward: inside the type test (:1472); backward: outside (:1616)
                            if (dir == 0 ? (tmatch && m->type == FORWARD) : (m->type == BACKWARD))
                                nk[i].push_back(Hinge{hv[i][k].pos, hv[i][k].type, false});
                        }
                    }
                }
            }
        }
    }
    // connected components, hinging.cpp:1644-1675
    {
        std::map<int, int> csize;
        for (int v = 0; v < num_hinges; v++) csize[find(v)]++;
        for (int i = 0; i < n_read; i++)
            for (size_t k = 0; k < hv[i].size(); k++)
                if (csize[find(node_base[i] + (int)k)] < p.min_cc_size) hv[i][k].active = false;
    }
    for (int i = 0; i < n_read; i++)  // hinging.cpp:1696-1704
        for (size_t j = 0; j < hv[i].size(); j++)
            if (active[i] && hv[i][j].active) {
                out->hinge_list.push_back(i);
                out->hinge_list.push_back(hinges_in[i][j].first);
                out->hinge_list.push_back(hinges_in[i][j].second);
            }

    // plain greedy, hinging.cpp:1724-1860
    for (int i = 0; i < n_read; i++) {
        if (!active[i]) continue;
        for (int dir = 0; dir < 2; dir++) {
            const std::vector<Ov*>& ms = dir == 0 ? mf[i] : mb[i];
            for (size_t j = 0; j < ms.size(); j++)
                if (ms[j]->active && ms[j]->type == (dir == 0 ? FORWARD : BACKWARD) &&
                    active[ms[j]->b]) {
                    out->greedy.push_back(make_edge(ms[j], -1));
                    break;
                }
        }
    }

    // best-overlap scoring loop, hinging.cpp:1911-2148
    int hinge_pos = -1;
    for (int i = 0; i < n_read; i++) {
        if (!active[i]) continue;
        for (int dir = 0; dir < 2; dir++) {
            const std::vector<Ov*>& ms = dir == 0 ? mf[i] : mb[i];
            const int plain = dir == 0 ? FORWARD : BACKWARD;
            const int internal = dir == 0 ? FORWARD_INTERNAL : BACKWARD_INTERNAL;
            int got = 0, got_internal = 0;
            const Ov* chosen = NULL;
            for (size_t j = 0; j < ms.size(); j++) {
                const Ov* m = ms[j];
                if (!m->active || !active[m->b]) continue;
                if (m->type == plain && got == 0) {
                    bool poisoned = false;
                    for (size_t k = 0; k < nk[i].size(); k++) {
                        bool hit;
                        if (dir == 0)
                            hit = (m->comp != 1 && nk[i][k].type == -1 && nk[i][k].pos > m->ebe) ||
                                  (m->comp == 1 && nk[i][k].type == 1 && nk[i][k].pos < m->ebs);
                        else
                            hit = (m->comp != 1 && nk[i][k].type == 1 && nk[i][k].pos < m->ebs) ||
                                  (m->comp == 1 && nk[i][k].type == -1 && nk[i][k].pos > m->ebe);
                        if (hit) {
                            out->skipped.push_back(make_edge(m, -1));
                            poisoned = true;
                        }
                    }
                    if (!poisoned) {
                        chosen = m;
                        hinge_pos = -1;
                        got = 1;
                    }
                } else if (m->type == internal && hv[m->b].size() > 0 && got_internal == 0) {
                    int bpos;
                    int want;
                    if (dir == 0) {
                        bpos = m->comp == 1 ? m->be : m->bs;
                        want = 1 - 2 * m->comp;
                    } else {
                        bpos = m->comp == 1 ? m->bs : m->be;
                        want = -1 + 2 * m->comp;
                    }
                    for (size_t k = 0; k < hv[m->b].size(); k++) {
                        const Hinge& h = hv[m->b][k];
                        if (bpos > h.pos - p.hinge_tolerance && bpos < h.pos + p.hinge_tolerance &&
                            h.type == want && h.active) {
                            if (got == 0 || m->weight > chosen->weight - 2 * p.hinge_slack) {
                                chosen = m;
                                got = 1;
                                got_internal = 1;
                                hinge_pos = h.pos;
                            }
                            break;
                        }
                    }
                }
            }
            if (chosen) {
                out->edges.push_back(make_edge(chosen, hinge_pos));
            } else {
                std::ostringstream ss;
                ss << i << "\t matches_" << (dir == 0 ? "forward" : "backward")
                   << " size: " << ms.size() << "\n";
                out->deadends.push_back(ss.str());
            }
        }
    }
}

// ------------------------------------------------------------------ text I/O

static void print_pairs(std::ofstream& f, int i, const std::vector<PII>& v) {
    f << i << " ";
    for (size_t j = 0; j < v.size(); j++) f << v[j].first << " " << v[j].second << " ";
    f << std::endl;
}

// filter.cpp:599-602,775-788,1078-1098
void write_filter_files(const FilterOut& o, const Params& p, int n_read, const std::string& x, int part) {
    (void)p;
    (void)n_read;
    const std::ios_base::openmode mode = part > 0 ? std::ios_base::app : std::ios_base::trunc;
    std::ofstream cov(x + ".coverage.txt", mode), homo(x + ".homologous.txt", mode), rep;
    if (part <= 0) rep.open(x + ".repeat.txt", mode);  // closed inside the part loop: part 0 only (filter.cpp:1086)
    std::ofstream filtered(x + ".filtered.fasta", mode), hg(x + ".hinges.txt", mode), mask(x + ".mas", mode);
    std::ofstream comask(x + ".cmas", mode), covflag(x + ".cov.flag", mode), selfflag(x + ".self.flag", mode);
    for (int i = o.r_begin; i <= o.r_end; i++) {
        cov << "read " << i << " ";
        for (size_t j = 0; j < o.cov0[i].size(); j++)
            cov << o.cov0[i][j].first << "," << o.cov0[i][j].second << " ";
        cov << std::endl;
        comask << i << " " << o.cmask[i].first << " " << o.cmask[i].second << std::endl;
        mask << i << " " << o.mask[i].first << " " << o.mask[i].second << std::endl;
        if (part <= 0) print_pairs(rep, i, o.repeats[i]);
        if (i < o.r_end) print_pairs(hg, i, o.hinges[i]);  // filter.cpp:1091 `i < r_end`
    }
    for (size_t k = 0; k < o.cov_flag.size(); k++) covflag << o.cov_flag[k] << std::endl;
    for (size_t k = 0; k < o.self_flag.size(); k++) selfflag << o.self_flag[k] << std::endl;
}

// maximal.cpp:853-878
void write_maximal_files(const MaximalOut& o, const std::vector<PII>& ranges, const std::string& x) {
    std::ofstream cont(x + ".contained.txt"), mx(x + ".max");
    for (size_t k = 0; k < o.contained.size(); k++)
        cont << o.contained[k].first << "\t" << o.contained[k].second << std::endl;
    for (const PII& r : ranges)
        for (int i = r.first; i <= r.second; i++)
            if (o.active[i]) mx << i << std::endl;
}

// hinging.cpp:188-248
static void print_edge(FILE* f, const Edge& e) {
    int hinged = (e.type == FORWARD || e.type == BACKWARD) ? -1 : 1;
    if (e.type == FORWARD_INTERNAL || e.type == FORWARD)
        fprintf(f, "%d %d %d %d %d %d [%d %d] [%d %d] [%d %d] [%d %d] [%d %d] [%d %d]\n", e.a, e.b,
                e.length, 0, e.comp, hinged, e.eas, e.eae, e.ebs, e.ebe, e.ras, e.rae, e.rbs, e.rbe,
                e.as, e.ae, e.bs, e.be);
    else
        fprintf(f, "%d %d %d %d %d %d [%d %d] [%d %d] [%d %d] [%d %d] [%d %d] [%d %d]\n", e.b, e.a,
                e.length, e.comp, 0, hinged, e.ebs, e.ebe, e.eas, e.eae, e.rbs, e.rbe, e.ras, e.rae,
                e.as, e.ae, e.bs, e.be);
}

// hinging.cpp:253-344
static void print_edge2(FILE* f, const Edge& e) {
    if (e.type == FORWARD || e.type == FORWARD_INTERNAL)
        fprintf(f, "%d %d %d %d %d %d %d [%d %d] [%d %d] [%d %d] [%d %d]\n", e.a, e.b, e.length, 0,
                e.comp, e.type == FORWARD ? 0 : 1, e.type == FORWARD ? -1 : e.hinge_pos, e.eas, e.eae,
                e.ebs, e.ebe, e.ras, e.rae, e.rbs, e.rbe);
    else
        fprintf(f, "%d %d %d %d %d %d %d [%d %d] [%d %d] [%d %d] [%d %d]\n", e.b, e.a, e.length,
                e.comp, 0, e.type == BACKWARD ? 0 : -1, e.type == BACKWARD ? -1 : e.hinge_pos, e.ebs,
                e.ebe, e.eas, e.eae, e.rbs, e.rbe, e.ras, e.rae);
}

void write_layout_files(const LayoutOut& o, int n_read, const std::string& x, const std::string& out) {
    {
        std::ofstream g(x + ".garbage.txt");
        for (size_t k = 0; k < o.garbage.size(); k++) g << o.garbage[k] << std::endl;
        std::ofstream dead(out + ".deadends.txt");
        for (size_t k = 0; k < o.deadends.size(); k++) dead << o.deadends[k];
        std::ofstream killed(x + ".killed.hinges");  // hinging.cpp:1201-1208
        for (int i = 0; i < n_read; i++) {
            killed << i << " ";
            for (size_t j = 0; j < o.killed[i].size(); j++)
                killed << o.killed[i][j].first << " " << o.killed[i][j].second << " ";
            killed << std::endl;
        }
    }
    FILE* f = fopen((out + ".hgraph").c_str(), "w");
    for (size_t k = 0; k < o.hgraph_lines.size(); k++) fputs(o.hgraph_lines[k].c_str(), f);
    fclose(f);
    f = fopen((out + ".hinge.list").c_str(), "w");
    for (size_t k = 0; k + 2 < o.hinge_list.size(); k += 3)
        fprintf(f, "%d %d %d\n", o.hinge_list[k], o.hinge_list[k + 1], o.hinge_list[k + 2]);
    fclose(f);
    f = fopen((out + ".edges.hinges").c_str(), "w");
    FILE* f2 = fopen((out + ".edges.hinges2").c_str(), "w");
    for (size_t k = 0; k < o.edges.size(); k++) {
        print_edge(f, o.edges[k]);
        print_edge2(f2, o.edges[k]);
    }
    fclose(f);
    fclose(f2);
    f = fopen((out + ".edges.skipped").c_str(), "w");
    for (size_t k = 0; k < o.skipped.size(); k++) print_edge(f, o.skipped[k]);
    fclose(f);
    f = fopen((out + ".edges.greedy").c_str(), "w");
    for (size_t k = 0; k < o.greedy.size(); k++) print_edge(f, o.greedy[k]);
    fclose(f);
    // .edges.1 / .edges.2: the same greedy edges in the older two-file format (hinging.cpp:1739-1786,
    // 1805-1852; forward and backward edges print the same way)
    f = fopen((out + ".edges.1").c_str(), "w");
    f2 = fopen((out + ".edges.2").c_str(), "w");
    for (size_t k = 0; k < o.greedy.size(); k++) {
        const Edge& e = o.greedy[k];
        fprintf(f, e.comp == 0 ? "%d %d %d [%d %d] [%d %d] [%d %d] [%d %d]\n" : "%d %d' %d [%d %d] [%d %d] [%d %d] [%d %d]\n",
                e.a, e.b, e.length, e.eas, e.eae, e.ebs, e.ebe, e.ras, e.rae, e.rbs, e.rbe);
        fprintf(f2, e.comp == 0 ? "%d' %d' %d [%d %d] [%d %d] [%d %d] [%d %d]\n" : "%d %d' %d [%d %d] [%d %d] [%d %d] [%d %d]\n",
                e.b, e.a, e.length, e.eas, e.eae, e.ebs, e.ebe, e.ras, e.rae, e.rbs, e.rbe);
    }
    fclose(f);
    fclose(f2);
}

void read_mask_file(const std::string& path, int n_read, std::vector<PII>* mask) {
    mask->assign(n_read, PII(0, 0));  // reference leaves absent reads uninitialised
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return;
    int r, s, e;
    while (fscanf(f, "%d %d %d", &r, &s, &e) == 3)
        if (r >= 0 && r < n_read) (*mask)[r] = PII(s, e);
    fclose(f);
}

void read_max_file(const std::string& path, int n_read, std::vector<char>* maximal) {
    maximal->assign(n_read, 0);
    std::ifstream f(path);
    std::string line;
    while (std::getline(f, line)) {
        int r = atoi(line.c_str());
        if (r >= 0 && r < n_read) (*maximal)[r] = 1;
    }
}

void read_pairs_file(const std::string& path, int n_read, std::vector<std::vector<PII>>* v) {
    v->assign(n_read, std::vector<PII>());
    std::ifstream f(path);
    std::string line;
    while (std::getline(f, line)) {
        std::stringstream ss;
        ss << line;
        int num = -1;
        ss >> num;
        if (num < 0 || num >= n_read) continue;
        (*v)[num].clear();
        while (!ss.eof()) {
            int r1 = 0, r2 = 0;
            ss >> r1 >> r2;
            if (r1 != 0 && r2 != 0) (*v)[num].push_back(PII(r1, r2));
        }
    }
}

}  // namespace oracle
