// hinge_synth — seeded generator of DAZZ_DB + .las fixtures (test/bench tooling).
//
// Produces what `daligner` would find on a random genome with planted exact
// repeats, without doing any alignment: every pair of reads that shares at
// least `min_ovl` bases of genome (or of two copies of a repeat family) yields
// one overlap record per direction, with coordinates from the shared interval,
// a few bases of end jitter, and a trace whose b-deltas add up to the B span.
// Record/stub/index/track layouts follow SURVEY.md Appendix B
// (/root/reference/src/include/DB.h:214-303, align.h:126-132,332-337), so the
// unmodified reference binaries open the files too.
//
// Deterministic: all randomness derives from `seed` through splitmix64, and the
// jitter of a pair is a symmetric hash of the two read ids, so the A->B and
// B->A records describe the same alignment and any A-range can be generated
// independently (used to shard the synthetic .las across GPUs).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

extern "C" {
typedef struct hgs_params {
    int64_t genome_len;
    double coverage;
    int32_t read_mean, read_sd, read_min, read_max;
    int32_t n_families;            // <0: genome_len / 250000
    int32_t rep_min_len, rep_max_len, rep_min_copies, rep_max_copies;
    int32_t min_ovl, jitter, tspace;
    double qv_bad_frac;
    uint64_t seed;
    double frag_prob;  // chance that a pair's alignment comes as two local alignments (as daligner
                       // reports long noisy overlaps: several records per (A,B) pair)
} hgs_params;
}

namespace {

inline uint64_t splitmix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() { return s = splitmix(s); }
    double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
    int64_t range(int64_t lo, int64_t hi) { return lo + (int64_t)(next() % (uint64_t)(hi - lo + 1)); }
    double normal() {
        double u1 = uni(), u2 = uni();
        if (u1 < 1e-300) u1 = 1e-300;
        return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    }
};

struct Copy {
    int64_t pos;
    int32_t len, family, idx;
};
struct CopyHit {  // a read's footprint on one repeat copy, in repeat coordinates
    int32_t x, y, read;
};

struct Rec {
    int32_t b, abpos, aepos, bbpos, bepos, flags, diffs;
};

struct Synth {
    hgs_params p;
    int32_t n_read = 0;
    std::vector<int64_t> gstart;
    std::vector<int32_t> rlen;
    std::vector<uint8_t> strand;
    std::vector<int32_t> by_start;  // read ids sorted by gstart
    std::vector<int64_t> sorted_start;
    int32_t max_rlen = 0;
    std::vector<Copy> copies;                  // sorted by pos
    std::vector<std::vector<int32_t>> family;  // family -> indices into copies
    std::vector<std::vector<CopyHit>> hits;    // per copy, sorted by x
    // generated block
    int32_t a_lo = 0, a_hi = 0;
    std::vector<int32_t> col[8];  // aread bread abpos aepos bbpos bepos diffs flags
    std::vector<int64_t> trace_off;
    std::vector<uint8_t> trace;
    std::vector<int64_t> qv_off;
    std::vector<uint8_t> qv;
};

void build(Synth* S) {
    const hgs_params& p = S->p;
    Rng rg(splitmix(p.seed ^ 0x1111));
    // repeat families: non-overlapping copies at random positions
    int nf = p.n_families < 0 ? (int)(p.genome_len / 250000) : p.n_families;
    std::vector<std::pair<int64_t, int64_t>> used;
    S->family.assign(nf, std::vector<int32_t>());
    for (int f = 0; f < nf; f++) {
        int len = (int)rg.range(p.rep_min_len, p.rep_max_len);
        int k = (int)rg.range(p.rep_min_copies, p.rep_max_copies);
        for (int c = 0; c < k; c++) {
            for (int attempt = 0; attempt < 100; attempt++) {
                if (p.genome_len <= len + 2) break;
                int64_t pos = rg.range(1, p.genome_len - len - 1);
                bool clash = false;
                for (auto& u : used)
                    if (pos < u.second + 2000 && u.first < pos + len + 2000) {
                        clash = true;
                        break;
                    }
                if (clash) continue;
                used.push_back(std::make_pair(pos, pos + len));
                Copy cp;
                cp.pos = pos; cp.len = len; cp.family = f; cp.idx = c;
                S->copies.push_back(cp);
                break;
            }
        }
    }
    std::sort(S->copies.begin(), S->copies.end(), [](const Copy& a, const Copy& b) { return a.pos < b.pos; });
    for (size_t i = 0; i < S->copies.size(); i++) S->family[S->copies[i].family].push_back((int32_t)i);

    // reads
    Rng rr(splitmix(p.seed ^ 0x2222));
    int64_t target = (int64_t)(p.coverage * (double)p.genome_len);
    int64_t total = 0;
    while (total < target) {
        int len = (int)lrint(p.read_mean + p.read_sd * rr.normal());
        if (len < p.read_min || len > p.read_max || len >= p.genome_len) continue;
        int64_t s = rr.range(0, p.genome_len - len);
        S->gstart.push_back(s);
        S->rlen.push_back(len);
        S->strand.push_back((uint8_t)(rr.next() & 1));
        total += len;
        if (len > S->max_rlen) S->max_rlen = len;
    }
    S->n_read = (int32_t)S->rlen.size();
    S->by_start.resize(S->n_read);
    for (int i = 0; i < S->n_read; i++) S->by_start[i] = i;
    std::sort(S->by_start.begin(), S->by_start.end(), [&](int a, int b) {
        return S->gstart[a] != S->gstart[b] ? S->gstart[a] < S->gstart[b] : a < b;
    });
    S->sorted_start.resize(S->n_read);
    for (int i = 0; i < S->n_read; i++) S->sorted_start[i] = S->gstart[S->by_start[i]];

    // footprints of reads on repeat copies
    S->hits.assign(S->copies.size(), std::vector<CopyHit>());
    for (size_t c = 0; c < S->copies.size(); c++) {
        const Copy& cp = S->copies[c];
        // reads with start < pos+len and end > pos
        size_t lo = std::lower_bound(S->sorted_start.begin(), S->sorted_start.end(), cp.pos - S->max_rlen) -
                    S->sorted_start.begin();
        for (size_t k = lo; k < (size_t)S->n_read && S->sorted_start[k] < cp.pos + cp.len; k++) {
            int r = S->by_start[k];
            int64_t s = std::max(S->gstart[r], cp.pos), e = std::min(S->gstart[r] + S->rlen[r], cp.pos + cp.len);
            if (e - s >= p.min_ovl) {
                CopyHit h;
                h.x = (int32_t)(s - cp.pos); h.y = (int32_t)(e - cp.pos); h.read = r;
                S->hits[c].push_back(h);
            }
        }
        std::sort(S->hits[c].begin(), S->hits[c].end(), [](const CopyHit& a, const CopyHit& b) {
            return a.x != b.x ? a.x < b.x : a.read < b.read;
        });
    }

    // intrinsic QVs: one byte per tspace tile
    Rng rq(splitmix(p.seed ^ 0x3333));
    S->qv_off.assign(S->n_read + 1, 0);
    for (int i = 0; i < S->n_read; i++) S->qv_off[i + 1] = S->qv_off[i] + (S->rlen[i] + p.tspace - 1) / p.tspace;
    S->qv.resize((size_t)S->qv_off[S->n_read]);
    for (size_t k = 0; k < S->qv.size(); k++) {
        uint64_t h = rq.next();
        bool bad = (h >> 11) * (1.0 / 9007199254740992.0) < p.qv_bad_frac;
        S->qv[k] = (uint8_t)(bad ? 40 + (h & 7) : 15 + (h & 15));
    }
}

// genome interval [gs,ge) -> read coordinates
inline void to_read(const Synth* S, int r, int64_t gs, int64_t ge, int* lo, int* hi) {
    int64_t s = S->gstart[r], e = s + S->rlen[r];
    if (!S->strand[r]) {
        *lo = (int)(gs - s);
        *hi = (int)(ge - s);
    } else {
        *lo = (int)(e - ge);
        *hi = (int)(e - gs);
    }
}

inline void emit(const Synth* S, int a, int b, int64_t gsa, int64_t gea, int64_t gsb, int64_t geb,
                 std::vector<Rec>* out) {
    int alo, ahi, blo, bhi;
    to_read(S, a, gsa, gea, &alo, &ahi);
    to_read(S, b, gsb, geb, &blo, &bhi);
    Rec r;
    r.b = b;
    r.abpos = alo;
    r.aepos = ahi;
    int comp = S->strand[a] ^ S->strand[b];
    if (!comp) {
        r.bbpos = blo;
        r.bepos = bhi;
    } else {
        r.bbpos = S->rlen[b] - bhi;
        r.bepos = S->rlen[b] - blo;
    }
    r.flags = comp;
    r.diffs = 0;
    out->push_back(r);
}

void overlaps_of(const Synth* S, int a, std::vector<Rec>* out) {
    const hgs_params& p = S->p;
    out->clear();
    const int64_t sa = S->gstart[a], ea = sa + S->rlen[a];
    // genomic neighbours
    size_t lo = std::lower_bound(S->sorted_start.begin(), S->sorted_start.end(), sa - S->max_rlen) -
                S->sorted_start.begin();
    for (size_t k = lo; k < (size_t)S->n_read && S->sorted_start[k] <= ea - p.min_ovl; k++) {
        int b = S->by_start[k];
        if (b == a) continue;
        int64_t sb = S->gstart[b], eb = sb + S->rlen[b];
        int64_t gs = std::max(sa, sb), ge = std::min(ea, eb);
        if (ge - gs < p.min_ovl) continue;
        uint64_t h = splitmix(p.seed ^ splitmix(((uint64_t)std::min(a, b) << 32) | (uint64_t)std::max(a, b)));
        int j1 = p.jitter ? (int)(h % (uint64_t)(p.jitter + 1)) : 0;
        int j2 = p.jitter ? (int)((h >> 20) % (uint64_t)(p.jitter + 1)) : 0;
        gs += j1;
        ge -= j2;
        if (ge - gs < p.min_ovl) continue;
        if (p.frag_prob > 0.0) {
            // split points and gaps come from the pair's hash, in genome coordinates: A->B and B->A
            // agree.  frag_prob in (1, 2]: a second split of the right-hand piece with chance frag_prob - 1.
            uint64_t h2 = h;
            double pr = p.frag_prob;
            for (int round = 0; round < 2 && pr > 0.0; round++, pr -= 1.0) {
                h2 = splitmix(h2 ^ 0x5151);
                const int gap = 50 + (int)((h2 >> 8) % 300);
                const int64_t room = ge - gs - 2 * (int64_t)p.min_ovl - gap;
                if (room < 0 || (h2 >> 40) * (1.0 / 16777216.0) >= pr) break;
                const int64_t cut = gs + p.min_ovl + (int64_t)((h2 >> 17) % (uint64_t)(room + 1));
                emit(S, a, b, gs, cut, gs, cut, out);
                gs = cut + gap;
            }
        }
        emit(S, a, b, gs, ge, gs, ge, out);
    }
    // repeat-mediated neighbours: A on copy c, B on another copy c2 of the family
    size_t c0 = std::lower_bound(S->copies.begin(), S->copies.end(), sa - (int64_t)p.rep_max_len,
                                 [](const Copy& c, int64_t v) { return c.pos < v; }) - S->copies.begin();
    for (size_t c = c0; c < S->copies.size() && S->copies[c].pos < ea; c++) {
        const Copy& cp = S->copies[c];
        int64_t s = std::max(sa, cp.pos), e = std::min(ea, cp.pos + cp.len);
        if (e - s < p.min_ovl) continue;
        int xa = (int)(s - cp.pos), ya = (int)(e - cp.pos);
        for (int32_t c2 : S->family[cp.family]) {
            if ((size_t)c2 == c) continue;
            const Copy& cq = S->copies[c2];
            const std::vector<CopyHit>& hs = S->hits[c2];
            CopyHit key;
            key.x = xa - cp.len; key.y = 0; key.read = -1;
            size_t k0 = std::lower_bound(hs.begin(), hs.end(), key, [](const CopyHit& u, const CopyHit& v) {
                return u.x < v.x; }) - hs.begin();
            for (size_t k = k0; k < hs.size() && hs[k].x <= ya - p.min_ovl; k++) {
                int b = hs[k].read;
                if (b == a) continue;
                int x = std::max(xa, hs[k].x), y = std::min(ya, hs[k].y);
                if (y - x < p.min_ovl) continue;
                uint64_t cc = ((uint64_t)std::min<int>((int)c, c2) << 20) ^ (uint64_t)std::max<int>((int)c, c2);
                uint64_t h = splitmix(p.seed ^ 0x7777 ^ splitmix(cc) ^
                                      splitmix(((uint64_t)std::min(a, b) << 32) | (uint64_t)std::max(a, b)));
                int j1 = p.jitter ? (int)(h % (uint64_t)(p.jitter + 1)) : 0;
                int j2 = p.jitter ? (int)((h >> 20) % (uint64_t)(p.jitter + 1)) : 0;
                x += j1;
                y -= j2;
                if (y - x < p.min_ovl) continue;
                emit(S, a, b, cp.pos + x, cp.pos + y, cq.pos + x, cq.pos + y, out);
            }
        }
    }
    std::sort(out->begin(), out->end(), [](const Rec& u, const Rec& v) {
        return u.b != v.b ? u.b < v.b : u.abpos < v.abpos;
    });
}

// trace of one record: (diff, bdelta) per tspace segment of A
inline int n_segments(const Rec& r, int tspace) { return (r.aepos - 1) / tspace - r.abpos / tspace + 1; }

int fill_trace(const Rec& r, int tspace, uint64_t h, uint8_t* t) {
    const int nseg = n_segments(r, tspace);
    const int blen = r.bepos - r.bbpos, alen = r.aepos - r.abpos;
    // segment lengths on A
    int prev = r.abpos, acc_b = 0, diffs = 0;
    for (int j = 0; j < nseg; j++) {
        int next = j == nseg - 1 ? r.aepos : (prev / tspace + 1) * tspace;
        int seg = next - prev;
        // proportional share of the B span, with a little zero-sum wobble
        int64_t target = (int64_t)(next - r.abpos) * blen / (alen > 0 ? alen : 1);
        int bd = (int)(target - acc_b);
        h = splitmix(h);
        if (j + 2 < nseg && seg == tspace) {
            int w = (int)(h % 7) - 3;
            if (bd + w >= 0 && bd + w <= 255) bd += w;
        }
        if (j == nseg - 1) bd = blen - acc_b;
        if (bd < 0) bd = 0;
        if (bd > 255) bd = 255;
        acc_b += bd;
        int df = (int)((h >> 8) % (uint64_t)(seg / 5 + 1));
        if (df > 255) df = 255;
        diffs += df;
        t[2 * j] = (uint8_t)df;
        t[2 * j + 1] = (uint8_t)bd;
        prev = next;
    }
    return diffs;
}

int64_t generate(Synth* S, int a_lo, int a_hi, int want_trace, int n_threads) {
    if (a_lo < 0) a_lo = 0;
    if (a_hi > S->n_read) a_hi = S->n_read;
    S->a_lo = a_lo;
    S->a_hi = a_hi;
    if (n_threads < 1) n_threads = 1;
    const int tspace = S->p.tspace;
    struct Part {
        std::vector<int32_t> col[8];
        std::vector<int64_t> tlen;
        std::vector<uint8_t> trace;
    };
    std::vector<Part> parts(n_threads);
    std::vector<std::thread> th;
    const int span = a_hi - a_lo;
    for (int t = 0; t < n_threads; t++) {
        th.emplace_back([&, t]() {
            int lo = a_lo + (int)((int64_t)span * t / n_threads);
            int hi = a_lo + (int)((int64_t)span * (t + 1) / n_threads);
            Part& P = parts[t];
            std::vector<Rec> recs;
            std::vector<uint8_t> tb;
            for (int a = lo; a < hi; a++) {
                overlaps_of(S, a, &recs);
                for (Rec& r : recs) {
                    int nseg = n_segments(r, tspace);
                    uint64_t h = splitmix(S->p.seed ^ splitmix(((uint64_t)a << 32) ^ (uint64_t)r.b) ^
                                          (uint64_t)r.abpos);
                    if (want_trace) {
                        tb.resize((size_t)2 * nseg);
                        r.diffs = fill_trace(r, tspace, h, tb.data());
                        P.trace.insert(P.trace.end(), tb.begin(), tb.end());
                    } else {
                        r.diffs = (r.aepos - r.abpos) / 8;
                    }
                    P.tlen.push_back(2 * nseg);
                    P.col[0].push_back(a); P.col[1].push_back(r.b); P.col[2].push_back(r.abpos);
                    P.col[3].push_back(r.aepos); P.col[4].push_back(r.bbpos); P.col[5].push_back(r.bepos);
                    P.col[6].push_back(r.diffs); P.col[7].push_back(r.flags);
                }
            }
        });
    }
    for (auto& t : th) t.join();
    int64_t n = 0, tbytes = 0;
    for (auto& P : parts) {
        n += (int64_t)P.col[0].size();
        tbytes += (int64_t)P.trace.size();
    }
    for (int c = 0; c < 8; c++) {
        S->col[c].clear();
        S->col[c].reserve((size_t)n + 8);
        for (auto& P : parts) {
            S->col[c].insert(S->col[c].end(), P.col[c].begin(), P.col[c].end());
            std::vector<int32_t>().swap(P.col[c]);
        }
    }
    S->trace_off.assign((size_t)n + 1, 0);
    S->trace.clear();
    S->trace.reserve((size_t)tbytes + 16);
    int64_t k = 0;
    for (auto& P : parts) {
        for (size_t i = 0; i < P.tlen.size(); i++, k++)
            S->trace_off[k + 1] = S->trace_off[k] + (want_trace ? P.tlen[i] : 0);
        S->trace.insert(S->trace.end(), P.trace.begin(), P.trace.end());
        std::vector<uint8_t>().swap(P.trace);
    }
    return n;
}

bool write_db(const Synth* S, const std::string& dir, const std::string& root, int with_bps, int with_qv) {
    const int n = S->n_read;
    // stub (DB.h:299-303); cutoff 0 / all 1 => the DB is never trimmed (DB.c:597)
    FILE* f = fopen((dir + "/" + root + ".db").c_str(), "w");
    if (!f) return false;
    fprintf(f, "files = %9d\n", 1);
    fprintf(f, "  %9d %s %s\n", n, root.c_str(), "Sim");
    fprintf(f, "blocks = %9d\n", 1);
    fprintf(f, "size = %9lld cutoff = %9d all = %1d\n", 400ll, 0, 1);
    fprintf(f, " %9d %9d\n", 0, 0);
    fprintf(f, " %9d %9d\n", n, n);
    fclose(f);
    // index: HITS_DB (112 B) then HITS_READ (40 B) per read (DB.h:214-291)
    f = fopen((dir + "/." + root + ".idx").c_str(), "wb");
    if (!f) return false;
    uint8_t hdr[112];
    memset(hdr, 0, sizeof hdr);
    int64_t totlen = 0;
    int maxlen = 0;
    for (int i = 0; i < n; i++) {
        totlen += S->rlen[i];
        maxlen = std::max(maxlen, S->rlen[i]);
    }
    int32_t iv[4] = {n, n, 0, 1};
    memcpy(hdr, iv, 16);
    float fr[4] = {0.25f, 0.25f, 0.25f, 0.25f};
    memcpy(hdr + 16, fr, 16);
    memcpy(hdr + 32, &maxlen, 4);
    memcpy(hdr + 40, &totlen, 8);
    fwrite(hdr, 1, sizeof hdr, f);
    int64_t boff = 0;
    for (int i = 0; i < n; i++) {
        uint8_t r[40];
        memset(r, 0, sizeof r);
        int32_t origin = i + 1, rl = S->rlen[i], fpulse = 0, flags = 0x0800 | 850;
        int64_t coff = -1;
        memcpy(r, &origin, 4);
        memcpy(r + 4, &rl, 4);
        memcpy(r + 8, &fpulse, 4);
        memcpy(r + 16, &boff, 8);
        memcpy(r + 24, &coff, 8);
        memcpy(r + 32, &flags, 4);
        fwrite(r, 1, sizeof r, f);
        boff += (rl + 3) >> 2;  // COMPRESSED_LEN, DB.h:193
    }
    fclose(f);
    if (with_bps) {  // bases are never inspected by the hot path: a sparse all-'a' file
        f = fopen((dir + "/." + root + ".bps").c_str(), "wb");
        if (!f) return false;
        fclose(f);
        if (truncate((dir + "/." + root + ".bps").c_str(), (off_t)boff) != 0) return false;
    }
    if (with_qv) {
        f = fopen((dir + "/." + root + ".qual.anno").c_str(), "wb");
        if (!f) return false;
        int32_t h2[2] = {n, 8};
        fwrite(h2, 4, 2, f);
        fwrite(S->qv_off.data(), 8, (size_t)n + 1, f);
        fclose(f);
        f = fopen((dir + "/." + root + ".qual.data").c_str(), "wb");
        if (!f) return false;
        fwrite(S->qv.data(), 1, S->qv.size(), f);
        fclose(f);
    }
    return true;
}

bool write_las(const Synth* S, const std::string& path) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    std::vector<char> iobuf(8 << 20);
    setvbuf(f, iobuf.data(), _IOFBF, iobuf.size());
    int64_t novl = (int64_t)S->col[0].size();
    int32_t tspace = S->p.tspace;
    fwrite(&novl, 8, 1, f);
    fwrite(&tspace, 4, 1, f);
    for (int64_t k = 0; k < novl; k++) {
        int64_t tb = S->trace_off[k + 1] - S->trace_off[k];
        int32_t rec[10] = {(int32_t)tb, S->col[6][k], S->col[2][k], S->col[4][k], S->col[3][k],
                           S->col[5][k], S->col[7][k], S->col[0][k], S->col[1][k], 0};
        fwrite(rec, 4, 10, f);
        if (tb) fwrite(&S->trace[S->trace_off[k]], 1, (size_t)tb, f);
    }
    fclose(f);
    return true;
}

}  // namespace

extern "C" {

void hgs_default_params(hgs_params* p) {
    p->genome_len = 1000000;
    p->coverage = 30.0;
    p->read_mean = 3500; p->read_sd = 1500; p->read_min = 1000; p->read_max = 60000;
    p->n_families = -1;
    p->rep_min_len = 3000; p->rep_max_len = 15000; p->rep_min_copies = 2; p->rep_max_copies = 4;
    p->min_ovl = 1000; p->jitter = 30; p->tspace = 100;
    p->qv_bad_frac = 0.002;
    p->seed = 1234;
    p->frag_prob = 0.0;
}

void* hgs_create(const hgs_params* p) {
    Synth* S = new Synth();
    S->p = *p;
    build(S);
    return S;
}
void hgs_destroy(void* h) { delete (Synth*)h; }
int32_t hgs_n_read(void* h) { return ((Synth*)h)->n_read; }
const int32_t* hgs_rlen(void* h) { return ((Synth*)h)->rlen.data(); }
const int64_t* hgs_qv_off(void* h) { return ((Synth*)h)->qv_off.data(); }
const uint8_t* hgs_qv(void* h) { return ((Synth*)h)->qv.data(); }
int32_t hgs_n_copies(void* h) { return (int32_t)((Synth*)h)->copies.size(); }
int64_t hgs_generate(void* h, int32_t a_lo, int32_t a_hi, int32_t want_trace, int32_t n_threads) {
    return generate((Synth*)h, a_lo, a_hi, want_trace, n_threads);
}
// col: 0 aread 1 bread 2 abpos 3 aepos 4 bbpos 5 bepos 6 diffs 7 flags
const int32_t* hgs_col(void* h, int32_t c) { return ((Synth*)h)->col[c].data(); }
const int64_t* hgs_trace_off(void* h) { return ((Synth*)h)->trace_off.data(); }
const uint8_t* hgs_trace(void* h) { return ((Synth*)h)->trace.data(); }
int64_t hgs_trace_bytes(void* h) { return (int64_t)((Synth*)h)->trace.size(); }
int32_t hgs_write_db(void* h, const char* dir, const char* root, int32_t with_bps, int32_t with_qv) {
    return write_db((Synth*)h, dir, root, with_bps, with_qv) ? 0 : -1;
}
int32_t hgs_write_las(void* h, const char* path) { return write_las((Synth*)h, path) ? 0 : -1; }

}  // extern "C"

#ifdef HGS_MAIN
int main(int argc, char** argv) {
    hgs_params p;
    hgs_default_params(&p);
    std::string dir = ".", root = "G";
    int threads = 8, bps = 1, qv = 1;
    for (int i = 1; i + 1 < argc; i += 2) {
        std::string k = argv[i];
        const char* v = argv[i + 1];
        if (k == "--genome") p.genome_len = atoll(v);
        else if (k == "--cov") p.coverage = atof(v);
        else if (k == "--read-mean") p.read_mean = atoi(v);
        else if (k == "--read-sd") p.read_sd = atoi(v);
        else if (k == "--read-min") p.read_min = atoi(v);
        else if (k == "--read-max") p.read_max = atoi(v);
        else if (k == "--families") p.n_families = atoi(v);
        else if (k == "--rep-min") p.rep_min_len = atoi(v);
        else if (k == "--rep-max") p.rep_max_len = atoi(v);
        else if (k == "--copies-min") p.rep_min_copies = atoi(v);
        else if (k == "--copies-max") p.rep_max_copies = atoi(v);
        else if (k == "--min-ovl") p.min_ovl = atoi(v);
        else if (k == "--jitter") p.jitter = atoi(v);
        else if (k == "--qv-bad") p.qv_bad_frac = atof(v);
        else if (k == "--seed") p.seed = strtoull(v, nullptr, 10);
        else if (k == "--frag") p.frag_prob = atof(v);
        else if (k == "--dir") dir = v;
        else if (k == "--root") root = v;
        else if (k == "--threads") threads = atoi(v);
        else if (k == "--bps") bps = atoi(v);
        else if (k == "--qv") qv = atoi(v);
        else {
            fprintf(stderr, "hinge_synth: unknown option %s\n", k.c_str());
            return 1;
        }
    }
    void* h = hgs_create(&p);
    int64_t n = hgs_generate(h, 0, hgs_n_read(h), 1, threads);
    mkdir(dir.c_str(), 0755);
    if (hgs_write_db(h, dir.c_str(), root.c_str(), bps, qv) != 0 ||
        hgs_write_las(h, (dir + "/" + root + ".las").c_str()) != 0) {
        fprintf(stderr, "hinge_synth: write failed\n");
        return 1;
    }
    printf("{\"n_read\": %d, \"novl\": %lld, \"repeat_copies\": %d}\n", hgs_n_read(h), (long long)n,
           hgs_n_copies(h));
    hgs_destroy(h);
    return 0;
}
#endif
