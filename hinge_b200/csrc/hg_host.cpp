// File-level drivers behind `hinge filter | maximal | layout`: same flags,
// inputs, outputs and exit codes as the reference executables
// (/root/reference/src/filter/filter.cpp:168-1123, maximal/maximal.cpp:238-905,
// layout/hinging.cpp:616-2156), with the compute done through the C ABI.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hinge_b200.h"
#include "hg_host.h"
#include "hg_io.h"

namespace hg {

// cmdline.h-style options: --name value | --name=value | -c value | --flag
bool parse_args(int argc, char** argv, bool layout, Args* a, std::string* err) {
    struct Opt {
        const char* name;
        char shortc;
        std::string* dst;
        bool* flag;
    };
    std::vector<Opt> opts = {
        {"db", 'b', &a->db, nullptr},         {"las", 'l', &a->las, nullptr},
        {"paf", 'p', &a->paf, nullptr},       {"config", 'c', &a->config, nullptr},
        {"fasta", 'f', &a->fasta, nullptr},   {"prefix", 'x', &a->prefix, nullptr},
        {"log", 'g', &a->log, nullptr},       {"mlas", 0, nullptr, &a->mlas},
        {"debug", 0, nullptr, &a->debug},
    };
    if (layout)
        opts.push_back({"out", 'o', &a->out, nullptr});
    else
        opts.push_back({"restrictreads", 'r', &a->restrictreads, nullptr});
    bool have_prefix = false, have_out = false;
    for (int i = 1; i < argc; i++) {
        std::string s = argv[i];
        const Opt* o = nullptr;
        std::string value;
        bool has_value = false;
        if (s.size() > 2 && s[0] == '-' && s[1] == '-') {
            std::string name = s.substr(2);
            size_t eq = name.find('=');
            if (eq != std::string::npos) {
                value = name.substr(eq + 1);
                name.resize(eq);
                has_value = true;
            }
            for (const Opt& c : opts)
                if (name == c.name) o = &c;
        } else if (s.size() == 2 && s[0] == '-') {
            for (const Opt& c : opts)
                if (c.shortc && s[1] == c.shortc) o = &c;
        }
        if (!o) {
            *err = "undefined option: " + s;
            return false;
        }
        if (o->flag) {
            *o->flag = true;
            continue;
        }
        if (!has_value) {
            if (i + 1 >= argc) {
                *err = std::string("option needs value: --") + o->name;
                return false;
            }
            value = argv[++i];
        }
        *o->dst = value;
        if (o->dst == &a->prefix) have_prefix = true;
        if (o->dst == &a->out) have_out = true;
    }
    if (layout && (!have_prefix || !have_out)) {  // hinging.cpp:627-628: required options
        *err = std::string("need option: --") + (!have_prefix ? "prefix" : "out");
        return false;
    }
    return true;
}

static void say(const char* fmt, const std::string& s = std::string()) {
    printf("[hinge_b200] ");
    printf(fmt, s.c_str());
    printf("\n");
    fflush(stdout);
}

// Flag-combination checks shared by the three stages (filter.cpp:212-241).
int check_inputs(const Args& a, std::vector<std::string>* las_names) {
    const bool db_and_las = !a.db.empty() && !a.las.empty();
    const bool db_or_las = !a.db.empty() || !a.las.empty();
    const bool fa_and_paf = !a.fasta.empty() && !a.paf.empty();
    const bool fa_or_paf = !a.fasta.empty() || !a.paf.empty();
    if (db_or_las && fa_or_paf) {
        fprintf(stderr, "Pass in either a db and a las or a fasta and a paf\n");
        return 1;
    }
    if (!fa_and_paf && !db_and_las) {
        fprintf(stderr, "Pass in at least one of the following two combinations: a db and a las or a fasta and a paf\n");
        return 1;
    }
    if (fa_and_paf) {
        fprintf(stderr, "hinge_b200: the fasta + paf input path is outside the B200 hot path (DESIGN.md, out of scope)\n");
        return 1;
    }
    las_names->clear();
    if (a.mlas) {  // the parts <las>.1.las, <las>.2.las, ... as far as they exist (filter.cpp:35-63)
        for (int i = 1;; i++) {
            const std::string name = a.las + "." + std::to_string(i) + ".las";
            struct stat st;
            if (stat(name.c_str(), &st) != 0) break;
            las_names->push_back(name);
        }
        if (las_names->empty()) {
            fprintf(stderr, "hinge_b200: --mlas: no file %s.1.las\n", a.las.c_str());
            return 1;
        }
        return 0;
    }
    std::string name = a.las;
    if (name.size() < 4 || name.compare(name.size() - 4, 4, ".las") != 0) name += ".las";
    las_names->push_back(name);
    return 0;
}

// Creating the CUDA context takes about half a second: it runs beside the reading of the inputs.
namespace {
struct EarlyContext {
    std::thread worker;
    hg_ctx* ctx = nullptr;
    int rc = HG_OK;
    bool started = false;
    double create_ms = 0;
    void start() {
        started = true;
        worker = std::thread([this]() {
            struct timespec a, b;
            clock_gettime(CLOCK_MONOTONIC, &a);
            rc = hg_ctx_create(0, nullptr, &ctx);
            clock_gettime(CLOCK_MONOTONIC, &b);
            create_ms = 1e3 * (double)(b.tv_sec - a.tv_sec) + 1e-6 * (double)(b.tv_nsec - a.tv_nsec);
        });
    }
    int take(hg_ctx** out) {
        if (!started) return hg_ctx_create(0, nullptr, out);
        if (worker.joinable()) worker.join();
        started = false;
        *out = ctx;
        ctx = nullptr;
        return rc;
    }
    void drop() {
        hg_ctx* c = nullptr;
        if (started && take(&c) == HG_OK) hg_ctx_destroy(c);
    }
} g_early;
}  // namespace

static int load_inputs_inner(const Args& a, bool want_trace, Ini* ini, ReadDB* db, LasFile* las,
                             std::vector<std::pair<int32_t, int32_t>>* part_ranges, bool load_las);

int load_inputs(const Args& a, bool want_trace, Ini* ini, ReadDB* db, LasFile* las,
                std::vector<std::pair<int32_t, int32_t>>* part_ranges, bool load_las) {
    std::vector<std::string> las_names;
    int rc = check_inputs(a, &las_names);
    if (rc) return rc;
    g_early.start();
    rc = load_inputs_inner(a, want_trace, ini, db, las, part_ranges, load_las);
    if (rc) g_early.drop();
    return rc;
}

static int load_inputs_inner(const Args& a, bool want_trace, Ini* ini, ReadDB* db, LasFile* las,
                             std::vector<std::pair<int32_t, int32_t>>* part_ranges, bool load_las) {
    std::vector<std::string> las_names;
    check_inputs(a, &las_names);
    if (db->open(a.db) != 0) {  // LAInterface::openDB exits 1
        fprintf(stderr, "%s\n", db->error.c_str());
        return 1;
    }
    say("# Reads: %s", std::to_string(db->n_read));
    if (ini->load(a.config) < 0) {  // filter.cpp:371-375
        fprintf(stderr, "Can't load %s\n", a.config.c_str());
        return 1;
    }
    if (!load_las) return 0;
    if (las->open_parts(las_names, want_trace, part_ranges) != 0) {
        fprintf(stderr, "%s\n", las->error.c_str());
        return 1;
    }
    say("# Alignments: %s", std::to_string(las->novl));
    if (las->novl == 0) {  // filter.cpp:505-508
        fprintf(stderr, "No alignments!\n");
        return 1;
    }
    return 0;
}

int open_context(const ReadDB& db, const LasFile& las, bool with_trace, hg_ctx** ctx) {
    PhaseTimer timer;
    int rc = g_early.take(ctx);
    timer.lap("  wait for CUDA context");
    timer.note("  (context creation, thread)", g_early.create_ms, "ms");
    if (rc != HG_OK) {
        fprintf(stderr, "hinge_b200: cannot create a CUDA context (status %d)\n", rc);
        return rc;
    }
    rc = hg_set_reads(*ctx, db.n_read, db.rlen.data(), db.has_qv ? db.qv_off.data() : nullptr,
                      db.has_qv ? db.qv.data() : nullptr, las.tspace);
    timer.lap("  hg_set_reads");
    if (rc == HG_OK)
        rc = hg_set_overlaps(*ctx, las.novl, las.aread.data(), las.bread.data(), las.abpos.data(),
                             las.aepos.data(), las.bbpos.data(), las.bepos.data(), las.diffs.data(),
                             las.flags.data(), with_trace ? las.trace_off.data() : nullptr,
                             with_trace ? las.trace.data() : nullptr, las.tbytes, HG_MEM_HOST, 0,
                             db.n_read);
    timer.lap("  hg_set_overlaps");
    if (rc != HG_OK) fprintf(stderr, "hinge_b200: %s\n", hg_last_error(*ctx));
    return rc;
}

void drop_early_context() { g_early.drop(); }

static bool g_exit_after = false;  // hg_main_exit_after
void release_context(hg_ctx* ctx) {
    if (!g_exit_after) hg_ctx_destroy(ctx);
}

static void touch(const std::string& path) {
    FILE* f = fopen(path.c_str(), "w");
    if (f) fclose(f);
}

}  // namespace hg

using namespace hg;

extern "C" void hg_main_exit_after(int on) { hg::g_exit_after = on != 0; }

extern "C" int hg_main_filter(int argc, char** argv) {
    mkdir("log", S_IRWXU | S_IRWXG | S_IROTH | S_IXOTH);  // filter.cpp:170
    Args a;
    std::string err;
    if (!parse_args(argc, argv, false, &a, &err)) {
        fprintf(stderr, "%s\n", err.c_str());
        return 1;
    }
    if (!a.restrictreads.empty()) {
        fprintf(stderr, "hinge_b200: --restrictreads (a debugging aid, filter.cpp:317-331) is not supported\n");
        return 1;
    }
    say("Reads filtering");
    PhaseTimer timer;
    Ini ini;
    ReadDB db;
    std::vector<std::string> las_names;
    if (check_inputs(a, &las_names)) return 1;
    {
        LasFile none;
        const int rc = load_inputs(a, false, &ini, &db, &none, nullptr, false);  // DB + INI; starts the CUDA context
        if (rc) return rc;
    }
    hg_filter_params fp;
    load_filter_params(ini, db.has_qv, &fp);
    const int n = db.n_read;
    const bool want_cov = getenv("HINGE_B200_SKIP_COVERAGE_TXT") == nullptr;
    const std::string& x = a.prefix;
    hg_ctx* ctx = nullptr;
    bool out_failed = false;
    double h2d_bytes = 0, d2h_bytes = 0;
    int64_t all_hinges = 0;

    // One pass per part.  A single .las is one part; with --mlas the reference loops over the parts and
    // carries state from one to the next (filter.cpp:474-1109), which makes its results differ from a
    // single-file run -- and this loop reproduces exactly that:
    //   * MIN_COV only ever grows: every part raises it to a third of ITS median coverage (:677-678)
    //   * the masks of the reads of earlier parts stay known; those of later parts are still (0,0)
    //     when a part calls its hinges (:534, 884-889)                         -> HG_OPT_KEEP_MASKS
    //   * .repeat.txt is closed after part 0 (:1086); every part's last read has no .hinges.txt line
    //     (:1091); all other files are appended to part after part
    for (size_t part = 0; part < las_names.size(); part++) {
        LasFile las;
        if (las.open(las_names[part], false) != 0) {
            fprintf(stderr, "%s\n", las.error.c_str());
            if (ctx) hg_ctx_destroy(ctx); else drop_early_context();
            return 1;
        }
        say("# Alignments: %s", std::to_string(las.novl));
        if (las.novl == 0) {  // filter.cpp:505-508
            fprintf(stderr, "No alignments!\n");
            if (ctx) hg_ctx_destroy(ctx); else drop_early_context();
            return 1;
        }
        timer.lap(part == 0 ? "read db + ini + las" : "read las part");
        int rc = HG_OK;
        if (part == 0) {
            rc = open_context(db, las, false, &ctx);
        } else {
            rc = hg_set_overlaps(ctx, las.novl, las.aread.data(), las.bread.data(), las.abpos.data(), las.aepos.data(),
                                 las.bbpos.data(), las.bepos.data(), las.diffs.data(), las.flags.data(), nullptr,
                                 nullptr, las.tbytes, HG_MEM_HOST, 0, n);
            if (rc != HG_OK) fprintf(stderr, "hinge_b200: %s\n", hg_last_error(ctx));
        }
        if (rc != HG_OK) {
            hg_ctx_destroy(ctx);
            return 1;
        }
        timer.lap("context + H2D + CSR");
        hg_set_option(ctx, HG_OPT_KEEP_COVERAGE, want_cov);
        hg_set_option(ctx, HG_OPT_KEEP_MASKS, part > 0);
        hg_filter_summary sum;
        rc = hg_filter(ctx, &fp, &sum);
        if (rc != HG_OK) {
            fprintf(stderr, "hinge_b200: filter failed: %s\n", hg_last_error(ctx));
            hg_ctx_destroy(ctx);
            return 1;
        }
        fp.min_cov = sum.min_cov;  // carried into the next part
        say("Estimated median coverage: %s", std::to_string(sum.cov_est));
        timer.lap("hg_filter");

        std::vector<int32_t> mask(2 * (size_t)n), cmask(2 * (size_t)n);
        std::vector<uint8_t> flags(n);
        std::vector<int64_t> anno_off((size_t)n + 1);
        std::vector<int32_t> apos((size_t)sum.n_annotations + 1), atype((size_t)sum.n_annotations + 1);
        std::vector<uint8_t> keep((size_t)sum.n_annotations + 1);
        rc = hg_filter_fetch(ctx, mask.data(), cmask.data(), flags.data(), anno_off.data(), apos.data(),
                             atype.data(), keep.data());
        std::vector<int64_t> cov_off;
        std::vector<int32_t> cov;
        if (rc == HG_OK && want_cov) {
            int64_t nb = 0;
            cov_off.resize((size_t)n + 1);
            rc = hg_filter_coverage(ctx, cov_off.data(), nullptr, &nb);
            cov.resize((size_t)nb + 1);
            if (rc == HG_OK) rc = hg_filter_coverage(ctx, nullptr, cov.data(), &nb);
        }
        if (rc != HG_OK) {
            fprintf(stderr, "hinge_b200: fetching results failed: %s\n", hg_last_error(ctx));
            hg_ctx_destroy(ctx);
            return 1;
        }
        const bool last_part = part + 1 == las_names.size();
        if (last_part) {
            release_context(ctx);   // the results are on the host
            ctx = nullptr;
        }
        timer.lap(last_part ? "fetch results + destroy" : "fetch results");
        h2d_bytes += 28.0 * (double)las.novl + (part == 0 ? 4.0 * n + (db.has_qv ? (double)db.qv.size() : 0.0) : 0.0);
        d2h_bytes += 17.0 * n + 8.0 * (n + 1) + 9.0 * (double)sum.n_annotations + 4.0 * (double)cov.size();

        {  // filter.cpp:449-457 opens all of these before the part loop, some stay empty
            const bool app = part > 0;
            if (!app) {
                touch(x + ".homologous.txt");
                touch(x + ".filtered.fasta");
                touch("debug.txt");
            }
            TextOut fmask(x + ".mas", app), fcmask(x + ".cmas", app);
            if (!want_cov && !app) touch(x + ".coverage.txt");
            TextOut fhg(x + ".hinges.txt", app), fcf(x + ".cov.flag", app), fsf(x + ".self.flag", app);
            TextOut frep(app ? std::string("/dev/null") : x + ".repeat.txt");  // closed after part 0 (filter.cpp:1086)
            int64_t hinges = 0;
            if (want_cov) {
                // .coverage.txt (filter.cpp:599-602) is ~11 bytes per 40-bp bin of every read, by far the
                // largest output: all cores format a block of reads at a time and write their pieces
                // straight to their places in the file (pwrite), no single-threaded copy in between
                const int fd = ::open((x + ".coverage.txt").c_str(), O_WRONLY | O_CREAT | (app ? 0 : O_TRUNC), 0644);
                if (fd < 0) {
                    fprintf(stderr, "hinge_b200: cannot write %s.coverage.txt\n", x.c_str());
                    out_failed = true;
                }
                off_t file_pos = fd >= 0 ? lseek(fd, 0, SEEK_END) : 0;
                int workers = (int)std::thread::hardware_concurrency();
                if (const char* v = getenv("HINGE_B200_IO_THREADS")) workers = atoi(v);
                workers = std::max(1, std::min(workers, 32));
                const int block = 4096 * workers;
                // raw buffers + a two-digits-at-a-time itoa: the text is ~11 bytes per bin and there are tens of
                // millions of bins, so the formatter is worth its own few lines
                std::vector<std::vector<char>> bufs((size_t)workers);
                static const char kDigits[] =
                    "00010203040506070809101112131415161718192021222324252627282930313233343536373839"
                    "40414243444546474849505152535455565758596061626364656667686970717273747576777879"
                    "8081828384858687888990919293949596979899";
                auto put_u = [](char* p, unsigned v) -> char* {  // v < 2^31
                    char tmp[12];
                    int n = 0;
                    while (v >= 100) {
                        const unsigned r = v % 100;
                        v /= 100;
                        tmp[n++] = kDigits[2 * r + 1];
                        tmp[n++] = kDigits[2 * r];
                    }
                    if (v >= 10) {
                        tmp[n++] = kDigits[2 * v + 1];
                        tmp[n++] = kDigits[2 * v];
                    } else {
                        tmp[n++] = (char)('0' + v);
                    }
                    while (n) *p++ = tmp[--n];
                    return p;
                };
                for (int b0 = sum.r_begin; b0 <= sum.r_end && fd >= 0; b0 += block) {
                    const int b1 = std::min(sum.r_end + 1, b0 + block);
                    std::vector<std::thread> pool;
                    for (int w = 0; w < workers; w++)
                        pool.emplace_back([&, w]() {
                            const int per = (b1 - b0 + workers - 1) / workers;
                            const int lo = std::min(b1, b0 + w * per), hi = std::min(b1, b0 + (w + 1) * per);
                            // "read <i> " + per bin "<pos>,<cov> " (at most 10 + 1 + 11 + 1 bytes) + newline
                            std::vector<char>& out = bufs[w];
                            out.resize((size_t)(hi - lo) * 32 + (size_t)(cov_off[hi] - cov_off[lo]) * 24 + 64);
                            char* p = out.data();
                            for (int i = lo; i < hi; i++) {
                                memcpy(p, "read ", 5);
                                p = put_u(p + 5, (unsigned)i);
                                *p++ = ' ';
                                unsigned pos = 0;
                                for (int64_t k = cov_off[i]; k < cov_off[i + 1]; k++, pos += (unsigned)fp.reso) {
                                    p = put_u(p, pos);
                                    *p++ = ',';
                                    int v = cov[k];
                                    if (v < 0) {
                                        *p++ = '-';
                                        v = -v;
                                    }
                                    p = put_u(p, (unsigned)v);
                                    *p++ = ' ';
                                }
                                *p++ = '\n';
                            }
                            out.resize((size_t)(p - out.data()));
                        });
                    for (auto& th : pool) th.join();
                    pool.clear();
                    std::vector<off_t> at((size_t)workers + 1, file_pos);
                    for (int w = 0; w < workers; w++) at[w + 1] = at[w] + (off_t)bufs[w].size();
                    // (tried: a shared mapping of the file's new tail filled by the same threads instead of pwrite --
                    // no faster, 0.49 vs 0.44 s for 640 MB: the cost is the page cache's, not the write path's)
                    for (int w = 0; w < workers; w++)
                        pool.emplace_back([&, w]() {
                            size_t done = 0;
                            while (done < bufs[w].size()) {
                                const ssize_t r = pwrite(fd, bufs[w].data() + done, bufs[w].size() - done, at[w] + (off_t)done);
                                if (r <= 0) break;
                                done += (size_t)r;
                            }
                        });
                    for (auto& th : pool) th.join();
                    file_pos = at[workers];
                }
                if (fd >= 0) ::close(fd);
            }
            for (int i = sum.r_begin; i <= sum.r_end; i++) {
                fcmask.put_int(i); fcmask.put_char(' '); fcmask.put_int(cmask[2 * i]); fcmask.put_char(' ');
                fcmask.put_int(cmask[2 * i + 1]); fcmask.put_char('\n');
                fmask.put_int(i); fmask.put_char(' '); fmask.put_int(mask[2 * i]); fmask.put_char(' ');
                fmask.put_int(mask[2 * i + 1]); fmask.put_char('\n');
                if (flags[i] & 1) { fcf.put_int(i); fcf.put_char('\n'); }
                if (flags[i] & 2) { fsf.put_int(i); fsf.put_char('\n'); }
                frep.put_int(i);  // filter.cpp:1078-1085
                frep.put_char(' ');
                for (int64_t k = anno_off[i]; k < anno_off[i + 1]; k++) {
                    frep.put_int(apos[k]); frep.put_char(' '); frep.put_int(atype[k]); frep.put_char(' ');
                }
                frep.put_char('\n');
                if (i < sum.r_end) {  // filter.cpp:1091: the last read (of every part) gets no line
                    fhg.put_int(i);
                    fhg.put_char(' ');
                    for (int64_t k = anno_off[i]; k < anno_off[i + 1]; k++)
                        if (keep[k]) {
                            fhg.put_int(apos[k]); fhg.put_char(' '); fhg.put_int(atype[k]); fhg.put_char(' ');
                            hinges++;
                        }
                    fhg.put_char('\n');
                }
            }
            out_failed = out_failed || !fmask.ok() || !fcmask.ok() || !fhg.ok() || !fcf.ok() || !fsf.ok() ||
                         !frep.ok();
            all_hinges += hinges;
            say("Number of hinges before filtering: %s", std::to_string(sum.n_annotations));
            say("Number of hinges: %s", std::to_string(hinges));
        }
        timer.lap("write output files");
    }
    timer.note("h2d bytes", h2d_bytes, "B");
    timer.note("d2h bytes", d2h_bytes, "B");
    return out_failed ? 1 : 0;
}
