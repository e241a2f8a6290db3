// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See hinge_oracle.h.
// ctypes-friendly array interface to the oracle, for parity tests on in-memory
// data and for bench.py's cpu_baseline ("port") leg.
#include <string.h>

#include "hinge_oracle.h"

using namespace oracle;

struct OrcHandle {
    Data d;
    Params p;
    FilterOut f;
    MaximalOut m;
    LayoutOut l;
    bool filter_done = false, maximal_done = false, layout_done = false;
};

extern "C" {

void* orc_create(int n_read, const int* rlen, const int64_t* qv_off, const uint8_t* qv, int tspace,
                 int64_t novl, const int* aread, const int* bread, const int* abpos, const int* aepos,
                 const int* bbpos, const int* bepos, const int* flags, const int64_t* trace_off,
                 const uint8_t* trace, int tbytes) {
    OrcHandle* h = new OrcHandle();
    Data& d = h->d;
    d.n_read = n_read;
    d.rlen.assign(rlen, rlen + n_read);
    d.has_qv = qv_off != nullptr && qv != nullptr;
    if (d.has_qv) {
        d.qv_off.assign(qv_off, qv_off + n_read + 1);
        d.qv.assign(qv, qv + qv_off[n_read]);
    }
    d.novl = novl;
    d.tspace = tspace;
    d.tbytes = tbytes;
    d.aread.assign(aread, aread + novl); d.bread.assign(bread, bread + novl);
    d.abpos.assign(abpos, abpos + novl); d.aepos.assign(aepos, aepos + novl);
    d.bbpos.assign(bbpos, bbpos + novl); d.bepos.assign(bepos, bepos + novl);
    d.flags.assign(flags, flags + novl);
    if (trace_off && trace) {
        d.trace_off.assign(trace_off, trace_off + novl + 1);
        d.trace.assign(trace, trace + trace_off[novl]);
    } else {
        d.trace_off.assign(novl + 1, 0);
    }
    return h;
}

void orc_destroy(void* h) { delete (OrcHandle*)h; }

int orc_load_ini(void* h, const char* path) {
    std::string err;
    return load_ini(path, &((OrcHandle*)h)->p, &err) ? 0 : -1;
}

// host threads for the loops with independent iterations (results do not depend on it)
void orc_set_threads(void* h, int threads) { ((OrcHandle*)h)->p.threads = threads < 1 ? 1 : threads; }

int orc_filter(void* hh) {
    OrcHandle* h = (OrcHandle*)hh;
    h->f = FilterOut();
    run_filter(h->d, h->p, &h->f);
    h->filter_done = true;
    return 0;
}

// sizes: mask/cmask 2*n_read, anno_off n_read+1, summary 4 (r_begin, r_end, cov_est, min_cov)
int64_t orc_filter_results(void* hh, int* mask, int* cmask, uint8_t* flags, int64_t* anno_off,
                           int* anno_pos, int* anno_type, uint8_t* hinge_keep, int* summary) {
    OrcHandle* h = (OrcHandle*)hh;
    if (!h->filter_done) return -1;
    const FilterOut& f = h->f;
    const int n = h->d.n_read;
    if (summary) {
        summary[0] = f.r_begin; summary[1] = f.r_end; summary[2] = f.cov_est; summary[3] = f.min_cov;
    }
    if (flags) {
        memset(flags, 0, n);
        for (int r : f.cov_flag) flags[r] |= 1;
        for (int r : f.self_flag) flags[r] |= 2;
    }
    int64_t o = 0;
    for (int i = 0; i < n; i++) {
        if (mask) { mask[2 * i] = f.mask[i].first; mask[2 * i + 1] = f.mask[i].second; }
        if (cmask) { cmask[2 * i] = f.cmask[i].first; cmask[2 * i + 1] = f.cmask[i].second; }
        if (anno_off) anno_off[i] = o;
        for (size_t k = 0; k < f.repeats[i].size(); k++, o++) {
            if (anno_pos) anno_pos[o] = f.repeats[i][k].first;
            if (anno_type) anno_type[o] = f.repeats[i][k].second;
            if (hinge_keep) {
                bool kept = false;
                for (const PII& hgp : f.hinges[i]) kept = kept || hgp == f.repeats[i][k];
                hinge_keep[o] = kept;
            }
        }
    }
    if (anno_off) anno_off[n] = o;
    return o;
}

}  // extern "C"
