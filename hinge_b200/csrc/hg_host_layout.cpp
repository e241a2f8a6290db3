// File-level drivers of `hinge maximal` and `hinge layout`
// (/root/reference/src/maximal/maximal.cpp:238-905, layout/hinging.cpp:616-2156).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/hinge_b200.h"
#include "hg_host.h"
#include "hg_io.h"
#include "hg_layout_result.h"

using namespace hg;

namespace {

void touch(const std::string& path) {
    FILE* f = fopen(path.c_str(), "w");
    if (f) fclose(f);
}

// maximal.cpp:524-531 / hinging.cpp:867-874: fscanf("%d %d %d") until EOF; reads
// without a line keep (0,0) here (the reference leaves them uninitialised)
bool read_mask_file(const std::string& path, int n_read, std::vector<int32_t>* mask) {
    mask->assign(2 * (size_t)n_read, 0);
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return false;
    int r, s, e;
    while (fscanf(f, "%d %d %d", &r, &s, &e) == 3)
        if (r >= 0 && r < n_read) {
            (*mask)[2 * (size_t)r] = s;
            (*mask)[2 * (size_t)r + 1] = e;
        }
    fclose(f);
    return true;
}

// hinging.cpp:398-412
void read_max_file(const std::string& path, int n_read, std::vector<uint8_t>* maximal) {
    maximal->assign(n_read, 0);
    std::ifstream f(path);
    std::string line;
    while (std::getline(f, line)) {
        const int r = atoi(line.c_str());
        if (r >= 0 && r < n_read) (*maximal)[r] = 1;
    }
}

// hinging.cpp:877-937: "<read> <pos> <type> <pos> <type> ...", pairs with a zero field are dropped
void read_pairs_file(const std::string& path, int n_read, std::vector<int64_t>* off,
                     std::vector<int32_t>* pos, std::vector<int32_t>* type) {
    std::vector<std::vector<std::pair<int, int>>> v(n_read);
    std::ifstream f(path);
    std::string line;
    while (std::getline(f, line)) {
        std::stringstream ss;
        ss << line;
        int num = -1;
        ss >> num;
        if (num < 0 || num >= n_read) continue;
        v[num].clear();
        while (!ss.eof()) {
            int r1 = 0, r2 = 0;
            ss >> r1 >> r2;
            if (r1 != 0 && r2 != 0) v[num].push_back(std::make_pair(r1, r2));
        }
    }
    off->assign((size_t)n_read + 1, 0);
    pos->clear();
    type->clear();
    for (int i = 0; i < n_read; i++) {
        for (auto& p : v[i]) {
            pos->push_back(p.first);
            type->push_back(p.second);
        }
        (*off)[i + 1] = (int64_t)pos->size();
    }
}

// hinging.cpp:188-248
void print_edge(FILE* f, const hg_edge& e) {
    const bool fwd = e.type == HG_FORWARD || e.type == HG_FORWARD_INTERNAL;
    const int hinged = (e.type == HG_FORWARD || e.type == HG_BACKWARD) ? -1 : 1;
    if (fwd)
        fprintf(f, "%d %d %d %d %d %d [%d %d] [%d %d] [%d %d] [%d %d] [%d %d] [%d %d]\n", e.a, e.b, e.length, 0,
                e.comp, hinged, e.eff_a[0], e.eff_a[1], e.eff_b[0], e.eff_b[1], e.read_a[0], e.read_a[1],
                e.read_b[0], e.read_b[1], e.raw_a[0], e.raw_a[1], e.raw_b[0], e.raw_b[1]);
    else
        fprintf(f, "%d %d %d %d %d %d [%d %d] [%d %d] [%d %d] [%d %d] [%d %d] [%d %d]\n", e.b, e.a, e.length,
                e.comp, 0, hinged, e.eff_b[0], e.eff_b[1], e.eff_a[0], e.eff_a[1], e.read_b[0], e.read_b[1],
                e.read_a[0], e.read_a[1], e.raw_a[0], e.raw_a[1], e.raw_b[0], e.raw_b[1]);
}

// hinging.cpp:253-344
void print_edge2(FILE* f, const hg_edge& e) {
    if (e.type == HG_FORWARD || e.type == HG_FORWARD_INTERNAL)
        fprintf(f, "%d %d %d %d %d %d %d [%d %d] [%d %d] [%d %d] [%d %d]\n", e.a, e.b, e.length, 0, e.comp,
                e.type == HG_FORWARD ? 0 : 1, e.type == HG_FORWARD ? -1 : e.hinge_pos, e.eff_a[0], e.eff_a[1],
                e.eff_b[0], e.eff_b[1], e.read_a[0], e.read_a[1], e.read_b[0], e.read_b[1]);
    else
        fprintf(f, "%d %d %d %d %d %d %d [%d %d] [%d %d] [%d %d] [%d %d]\n", e.b, e.a, e.length, e.comp, 0,
                e.type == HG_BACKWARD ? 0 : -1, e.type == HG_BACKWARD ? -1 : e.hinge_pos, e.eff_b[0],
                e.eff_b[1], e.eff_a[0], e.eff_a[1], e.read_b[0], e.read_b[1], e.read_a[0], e.read_a[1]);
}

// the 13-integer debug line of edges.g_out.txt / edges.*.backup.txt (hinging.cpp:1080-1090)
void print_match_debug(FILE* f, const hg_edge& e) {
    fprintf(f, "%d %d %d %d %d [%d %d] [%d %d] [%d %d] [%d %d] \n", e.a, e.b, e.length, e.comp, e.type,
            e.eff_a[0], e.eff_a[1], e.eff_b[0], e.eff_b[1], e.read_a[0], e.read_a[1], e.read_b[0], e.read_b[1]);
}

// .edges.1 / .edges.2 (hinging.cpp:1739-1786)
void print_greedy12(FILE* g1, FILE* g2, const hg_edge& e) {
    fprintf(g1, e.comp == 0 ? "%d %d %d [%d %d] [%d %d] [%d %d] [%d %d]\n" : "%d %d' %d [%d %d] [%d %d] [%d %d] [%d %d]\n",
            e.a, e.b, e.length, e.eff_a[0], e.eff_a[1], e.eff_b[0], e.eff_b[1], e.read_a[0], e.read_a[1],
            e.read_b[0], e.read_b[1]);
    fprintf(g2, e.comp == 0 ? "%d' %d' %d [%d %d] [%d %d] [%d %d] [%d %d]\n" : "%d %d' %d [%d %d] [%d %d] [%d %d] [%d %d]\n",
            e.b, e.a, e.length, e.eff_a[0], e.eff_a[1], e.eff_b[0], e.eff_b[1], e.read_a[0], e.read_a[1],
            e.read_b[0], e.read_b[1]);
}

// Output files: an unwritable path must end in exit code 1, not in fprintf(NULL).
bool g_out_failed = false;
FILE* open_out(const std::string& path) {
    FILE* f = fopen(path.c_str(), "w");
    if (!f) {
        fprintf(stderr, "hinge_b200: cannot write %s\n", path.c_str());
        g_out_failed = true;
        f = fopen("/dev/null", "w");
    }
    return f;
}

}  // namespace

extern "C" int hg_main_maximal(int argc, char** argv) {
    mkdir("log", S_IRWXU | S_IRWXG | S_IROTH | S_IXOTH);
    Args a;
    std::string err;
    if (!parse_args(argc, argv, false, &a, &err)) {
        fprintf(stderr, "%s\n", err.c_str());
        return 1;
    }
    PhaseTimer timer;
    Ini ini;
    ReadDB db;
    LasFile las;
    // --mlas: the parts hold disjoint, ascending A-read ranges and the reference's part loop only shares
    // the per-read active flags (maximal.cpp:562-900), so the records are taken together; the part
    // ranges still decide which reads .max lists
    std::vector<std::pair<int32_t, int32_t>> part_ranges;
    int rc = load_inputs(a, true, &ini, &db, &las, &part_ranges);
    if (rc) return rc;
    hg_layout_params lp;
    load_layout_params(ini, &lp);
    const int n = db.n_read;
    std::vector<int32_t> mask;
    if (!read_mask_file(a.prefix + ".mas", n, &mask)) {
        fprintf(stderr, "hinge maximal: cannot read %s.mas (run hinge filter first)\n", a.prefix.c_str());
        drop_early_context();
        return 1;
    }
    timer.lap("read db + ini + las + mas");
    hg_ctx* ctx = nullptr;
    if (open_context(db, las, true, &ctx) != HG_OK) {
        hg_ctx_destroy(ctx);
        return 1;
    }
    timer.lap("context + H2D + CSR");
    std::vector<uint8_t> maximal(n);
    std::vector<int32_t> by(n, -1);
    const bool want_contained = getenv("HINGE_B200_SKIP_CONTAINED_TXT") == nullptr;
    float ms = 0;
    rc = hg_maximal(ctx, &lp, mask.data(), maximal.data(), want_contained ? by.data() : nullptr, &ms);
    if (rc != HG_OK) {
        fprintf(stderr, "hinge_b200: maximal failed: %s\n", hg_last_error(ctx));
        hg_ctx_destroy(ctx);
        return 1;
    }
    timer.lap("hg_maximal");
    release_context(ctx);
    touch(a.prefix + ".homologous.txt");  // maximal.cpp:515-517 reopens (truncates) these
    touch(a.prefix + ".filtered.fasta");
    TextOut fmax(a.prefix + ".max"), fcont(a.prefix + ".contained.txt");
    int kept = 0;
    for (const auto& pr : part_ranges)
        for (int i = pr.first; i <= pr.second; i++) {
            if (by[i] >= 0) {  // maximal.cpp:853-857
                fcont.put_int(i);
                fcont.put_char('\t');
                fcont.put_int(by[i]);
                fcont.put_char('\n');
            }
            if (maximal[i]) {  // maximal.cpp:873-878
                fmax.put_int(i);
                fmax.put_char('\n');
                kept++;
            }
        }
    printf("[hinge_b200] removed contained reads, active reads: %d (%.3f ms on device)\n", kept, ms);
    timer.lap("destroy + write output files");
    return 0;
}

extern "C" int hg_main_layout(int argc, char** argv) {
    g_out_failed = false;
    mkdir("log", S_IRWXU | S_IRWXG | S_IROTH | S_IXOTH);
    Args a;
    std::string err;
    if (!parse_args(argc, argv, true, &a, &err)) {
        fprintf(stderr, "%s\n", err.c_str());
        return 1;
    }
    printf("[hinge_b200] Hinging layout\n");
    PhaseTimer timer;
    Ini ini;
    ReadDB db;
    LasFile las;
    int rc = load_inputs(a, true, &ini, &db, &las);
    if (rc) return rc;
    hg_layout_params lp;
    load_layout_params(ini, &lp);
    const int n = db.n_read;
    const std::string &x = a.prefix, &o = a.out;
    std::vector<int32_t> mask;
    std::vector<uint8_t> maximal;
    std::vector<int64_t> rep_off, hin_off;
    std::vector<int32_t> rep_pos, rep_type, hin_pos, hin_type;
    if (!read_mask_file(x + ".mas", n, &mask)) {
        fprintf(stderr, "hinge layout: cannot read %s.mas (run hinge filter first)\n", x.c_str());
        drop_early_context();
        return 1;
    }
    read_max_file(x + ".max", n, &maximal);
    read_pairs_file(x + ".repeat.txt", n, &rep_off, &rep_pos, &rep_type);
    read_pairs_file(x + ".hinges.txt", n, &hin_off, &hin_pos, &hin_type);
    // pointers must be valid even for empty lists
    rep_pos.push_back(0); rep_type.push_back(0); hin_pos.push_back(0); hin_type.push_back(0);
    timer.lap("read inputs");

    hg_ctx* ctx = nullptr;
    if (open_context(db, las, true, &ctx) != HG_OK) {
        hg_ctx_destroy(ctx);
        return 1;
    }
    timer.lap("context + H2D + CSR");
    float ms = 0;
    rc = hg_layout(ctx, &lp, mask.data(), maximal.data(), rep_off.data(), rep_pos.data(), rep_type.data(),
                   hin_off.data(), hin_pos.data(), hin_type.data(), &ms);
    if (rc != HG_OK) {
        fprintf(stderr, "hinge_b200: layout failed: %s\n", hg_last_error(ctx));
        hg_ctx_destroy(ctx);
        return 1;
    }
    const LayoutResult& R = *layout_result(ctx);
    timer.lap("hg_layout");

    {  // files that only depend on the inputs and the hinge bookkeeping
        TextOut garbage(x + ".garbage.txt");  // hinging.cpp:954-960
        for (int r : R.garbage) {
            garbage.put_int(r);
            garbage.put_char('\n');
        }
        TextOut killed(x + ".killed.hinges");  // hinging.cpp:1201-1208: "<read> <type> <pos> ..."
        for (int i = 0; i < n; i++) {
            killed.put_int(i);
            killed.put_char(' ');
            for (int64_t k = R.kil_off[i]; k < R.kil_off[i + 1]; k++) {
                killed.put_int(R.kil_type[k]); killed.put_char(' ');
                killed.put_int(R.kil_pos[k]); killed.put_char(' ');
            }
            killed.put_char('\n');
        }
        FILE* f = open_out(o + ".hgraph");  // hinging.cpp:1421-1626
        for (const GraphRec& g : R.graph)
            fprintf(f, "%d %d %d %d %d %d\n", g.f[0], g.f[1], g.f[2], g.f[3], g.flag, g.rev);
        fclose(f);
        f = open_out(o + ".hinge.list");  // hinging.cpp:1696-1704
        for (int i = 0; i < n; i++)
            for (int64_t k = R.hin_off[i]; k < R.hin_off[i + 1]; k++)
                if (R.active[i] && R.hin_alive[k]) fprintf(f, "%d %d %d\n", i, R.hin_pos[k], R.hin_type[k]);
        fclose(f);
        touch(o + ".debug");  // only written for a case the reference calls impossible (hinging.cpp:1474-1497)
        touch("overlap_debug.txt");
        touch("hinge_debug.txt");
    }

    hg_edge e;
    {  // debugging dumps in the working directory (hinging.cpp:1073-1151)
        FILE* g = open_out("edges.g_out.txt");
        FILE* fb = open_out("edges.fwd.backup.txt");
        FILE* bb = open_out("edges.bkw.backup.txt");
        for (int half = 0; half < 2; half++) {
            if (half) fprintf(g, "bkw\n");
            for (int i = 0; i < n; i++) {
                if (!R.active[i]) continue;
                const int lo = half ? R.ranges[i].z : R.ranges[i].x, hi = half ? R.ranges[i].w : R.ranges[i].y;
                bool first = true;
                for (int t = lo; t < hi; t++) {
                    const int ci = R.order[t];
                    if (!R.active[R.cands[ci].b]) continue;
                    R.fill_edge(ci, -1, &e);
                    if (first) print_match_debug(g, e);
                    first = false;
                    print_match_debug(half ? bb : fb, e);
                }
            }
        }
        fclose(g);
        fclose(fb);
        fclose(bb);
    }
    {  // plain greedy graph (hinging.cpp:1724-1860)
        FILE* g1 = open_out(o + ".edges.1");
        FILE* g2 = open_out(o + ".edges.2");
        FILE* gr = open_out(o + ".edges.greedy");
        for (int i = 0; i < n; i++) {
            if (!R.active[i]) continue;
            for (int half = 0; half < 2; half++) {
                const int lo = half ? R.ranges[i].z : R.ranges[i].x, hi = half ? R.ranges[i].w : R.ranges[i].y;
                for (int t = lo; t < hi; t++) {
                    const Cand& c = R.cands[R.order[t]];
                    if (c.type == (half ? HG_BACKWARD : HG_FORWARD) && R.active[c.b]) {
                        R.fill_edge(R.order[t], -1, &e);
                        print_edge(gr, e);
                        print_greedy12(g1, g2, e);
                        break;
                    }
                }
            }
        }
        fclose(g1);
        fclose(g2);
        fclose(gr);
    }
    {  // the hinge-aware graph (hinging.cpp:1911-2148)
        FILE* hg = open_out(o + ".edges.hinges");
        FILE* hg2 = open_out(o + ".edges.hinges2");
        FILE* sk = open_out(o + ".edges.skipped");
        std::ofstream dead(o + ".deadends.txt");
        for (const SkipRec& s : R.skips) {
            R.fill_edge(s.cand, -1, &e);
            print_edge(sk, e);
        }
        int64_t n_edges = 0;
        for (int i = 0; i < n; i++) {
            if (!R.active[i]) continue;
            for (int half = 0; half < 2; half++) {
                const int2 ch = R.chosen[2 * (size_t)i + half];
                if (ch.x >= 0) {
                    R.fill_edge(ch.x, ch.y, &e);
                    print_edge(hg, e);
                    print_edge2(hg2, e);
                    n_edges++;
                } else {
                    const int sz = half ? R.ranges[i].w - R.ranges[i].z : R.ranges[i].y - R.ranges[i].x;
                    dead << i << "\t matches_" << (half ? "backward" : "forward") << " size: " << sz << std::endl;
                }
            }
        }
        fclose(hg);
        fclose(hg2);
        fclose(sk);
        printf("[hinge_b200] %lld edges, %.3f ms on device\n", (long long)n_edges, ms);
    }
    release_context(ctx);
    timer.lap("write output files + destroy");
    return g_out_failed ? 1 : 0;
}
