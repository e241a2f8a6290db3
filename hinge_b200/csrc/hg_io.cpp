// See hg_io.h.  Written from the on-disk formats, not from the reference's
// reader code.
#include "hg_io.h"

#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <thread>

namespace hg {

namespace {

const int kDbBest = 0x0800;  // DB.h:211
const size_t kHitsDbBytes = 112;   // sizeof(HITS_DB) on LP64 (DB.h:259-291)
const size_t kHitsReadBytes = 40;  // sizeof(HITS_READ) (DB.h:214-222)

bool read_file(const std::string& path, std::vector<uint8_t>* out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseeko(f, 0, SEEK_END);
    off_t n = ftello(f);
    fseeko(f, 0, SEEK_SET);
    out->resize((size_t)n);
    bool ok = n == 0 || fread(out->data(), 1, (size_t)n, f) == (size_t)n;
    fclose(f);
    return ok;
}

void split_path(const std::string& name, std::string* dir, std::string* root) {
    std::string s = name;
    if (s.size() > 3 && s.compare(s.size() - 3, 3, ".db") == 0) s.resize(s.size() - 3);
    size_t slash = s.rfind('/');
    if (slash == std::string::npos) {
        *dir = ".";
        *root = s;
    } else {
        *dir = slash == 0 ? "/" : s.substr(0, slash);
        *root = s.substr(slash + 1);
    }
}

}  // namespace

int ReadDB::open(const std::string& db_name) {
    std::string dir, root;
    split_path(db_name, &dir, &root);

    // --- stub: only cutoff/all matter here (DB.h:299-303, DB.c:452-497)
    int cutoff = 0, all = 1;
    {
        FILE* f = fopen((dir + "/" + root + ".db").c_str(), "r");
        if (!f) {
            error = "Could not open database " + db_name;
            return -1;
        }
        int nfiles = 0;
        if (fscanf(f, "files = %9d\n", &nfiles) != 1) {
            fclose(f);
            error = "Stub file (.db) of " + root + " is junk";
            return -1;
        }
        for (int i = 0; i < nfiles; i++) {
            int last;
            char a[10000], b[10000];
            if (fscanf(f, "  %9d %9999s %9999s\n", &last, a, b) != 3) {
                fclose(f);
                error = "Stub file (.db) of " + root + " is junk";
                return -1;
            }
        }
        int nblocks = 0;
        if (fscanf(f, "blocks = %9d\n", &nblocks) == 1) {
            long long size;
            if (fscanf(f, "size = %9lld cutoff = %9d all = %1d\n", &size, &cutoff, &all) != 3) {
                fclose(f);
                error = "Stub file (.db) of " + root + " is junk";
                return -1;
            }
        }
        fclose(f);
    }

    // --- index
    std::vector<uint8_t> idx;
    if (!read_file(dir + "/." + root + ".idx", &idx) || idx.size() < kHitsDbBytes) {
        error = "Index file (.idx) of " + root + " is junk";
        return -1;
    }
    int32_t ureads, treads;
    memcpy(&ureads, idx.data(), 4);
    memcpy(&treads, idx.data() + 4, 4);
    if (ureads < 0 || idx.size() < kHitsDbBytes + kHitsReadBytes * (size_t)ureads) {
        error = "Index file (.idx) of " + root + " is junk";
        return -1;
    }
    // Trim_DB (DB.c:585-683): ids used by the .las are positions among kept reads
    const bool trimmed = !(cutoff <= 0 && all);
    const int allflag = all ? 0 : kDbBest;
    std::vector<int32_t> kept;  // untrimmed index of every kept read
    rlen.clear();
    for (int32_t i = 0; i < ureads; i++) {
        const uint8_t* r = idx.data() + kHitsDbBytes + kHitsReadBytes * (size_t)i;
        int32_t len, fl;
        memcpy(&len, r + 4, 4);
        memcpy(&fl, r + 32, 4);
        if (!trimmed || ((fl & kDbBest) >= allflag && len >= cutoff)) {
            kept.push_back(i);
            rlen.push_back(len);
        }
    }
    n_read = (int32_t)rlen.size();

    // --- qual track (DB.c:1080-1323): optional
    has_qv = false;
    qv_off.clear();
    qv.clear();
    std::vector<uint8_t> anno, data;
    if (read_file(dir + "/." + root + ".qual.anno", &anno) && anno.size() >= 8) {
        int32_t tracklen, size;
        memcpy(&tracklen, anno.data(), 4);
        memcpy(&size, anno.data() + 4, 4);
        if (size == 0) size = 8;
        const bool for_untrimmed = tracklen == ureads;
        const bool for_trimmed = !for_untrimmed && tracklen == treads && tracklen == n_read;
        if ((size == 8 || size == 4) && (for_untrimmed || for_trimmed) &&
            anno.size() >= 8 + (size_t)size * ((size_t)tracklen + 1) &&
            read_file(dir + "/." + root + ".qual.data", &data)) {
            auto off_at = [&](int64_t k) -> int64_t {
                if (size == 8) {
                    int64_t v;
                    memcpy(&v, anno.data() + 8 + 8 * k, 8);
                    return v;
                }
                int32_t v;
                memcpy(&v, anno.data() + 8 + 4 * k, 4);
                return v;
            };
            qv_off.assign((size_t)n_read + 1, 0);
            bool ok = true;
            for (int32_t j = 0; j < n_read && ok; j++) {
                int64_t src = for_untrimmed ? kept[j] : j;
                int64_t b = off_at(src), e = off_at(src + 1);
                if (b < 0 || e < b || (size_t)e > data.size()) ok = false;
                qv_off[j + 1] = qv_off[j] + (e - b);
            }
            if (ok) {
                qv.resize((size_t)qv_off[n_read]);
                for (int32_t j = 0; j < n_read; j++) {
                    int64_t src = for_untrimmed ? kept[j] : j;
                    int64_t b = off_at(src);
                    memcpy(qv.data() + qv_off[j], data.data() + b, (size_t)(qv_off[j + 1] - qv_off[j]));
                }
                has_qv = true;
            } else {
                qv_off.clear();
            }
        }
    }
    return 0;
}

// A .las is a chain: record i + 1 starts where the trace of record i ends, so finding the
// records is a pointer chase over the file.  The ingest runs the chase on T chunks at once:
// every worker but the first guesses where a record starts inside its chunk (the first offset
// from which kSyncChain records in a row look sane: lengths, coordinates and A-read order, and
// the trace length daligner writes for them, align.c:3124-3148), walks to the end of its chunk
// and the walks are then CHECKED to join up exactly -- worker t must end on worker t + 1's
// guess.  If any joint is off (a file that breaks the heuristics), the whole thing is redone
// by one sequential walk, so the result never depends on the guess.
namespace {

const size_t kRec = 40;  // tlen diffs abpos bbpos aepos bepos flags aread bread pad (align.h:126-132,332-337)
const int kSyncChain = 8;

struct RecHead {
    int32_t tlen, diffs, abpos, bbpos, aepos, bepos, flags, aread, bread;
};

inline bool plausible(const RecHead& r, int tspace) {
    if (r.tlen < 0 || (r.tlen & 1) || r.diffs < 0 || r.abpos < 0 || r.aepos <= r.abpos || r.bbpos < 0 ||
        r.bepos < r.bbpos || r.aread < 0 || r.bread < 0 || (r.flags & ~0xff))
        return false;
    return r.tlen == 2 * ((r.aepos - 1) / tspace - r.abpos / tspace + 1);
}

// First offset >= from (and < limit) where a chain of plausible records starts; fsize if none.
size_t find_sync(const uint8_t* base, size_t fsize, size_t from, size_t limit, int tspace, int tbytes) {
    // record offsets are multiples of 2 * tbytes: 12-byte header, 40-byte records, even trace lengths
    const size_t step = 2 * (size_t)tbytes;
    for (size_t p = (from + step - 1) / step * step; p < limit && p + kRec <= fsize; p += step) {
        size_t q = p;
        int prev_a = -1, k = 0;
        for (; k < kSyncChain; k++) {
            if (q == fsize) break;  // chain runs into the end of the file: fine
            if (q + kRec > fsize) {
                k = -1;
                break;
            }
            RecHead r;
            memcpy(&r, base + q, sizeof r);
            if (!plausible(r, tspace) || r.aread < prev_a) {
                k = -1;
                break;
            }
            prev_a = r.aread;
            q += kRec + (size_t)r.tlen * tbytes;
            if (q > fsize) {
                k = -1;
                break;
            }
        }
        if (k >= 0) return p;
    }
    return fsize;
}

struct Walk {
    size_t begin = 0, end = 0;       // byte range actually walked: [begin, end)
    std::vector<uint64_t> rec_off;   // start of every record found
    int64_t trace_bytes = 0;
    bool ok = true;
};

// Walks the records from `begin` until the first record start >= stop.
void walk_records(const uint8_t* base, size_t fsize, size_t begin, size_t stop, int tbytes, Walk* w) {
    w->begin = begin;
    size_t p = begin;
    while (p < stop) {
        if (p + kRec > fsize) {
            w->ok = false;
            break;
        }
        int32_t tlen;
        memcpy(&tlen, base + p, 4);
        const size_t tb = (size_t)tlen * (size_t)tbytes;
        if (tlen < 0 || p + kRec + tb > fsize) {
            w->ok = false;
            break;
        }
        w->rec_off.push_back(p);
        w->trace_bytes += (int64_t)tb;
        p += kRec + tb;
    }
    w->end = p;
}

}  // namespace

int LasFile::open_parts(const std::vector<std::string>& names, bool want_trace,
                        std::vector<std::pair<int32_t, int32_t>>* ranges) {
    if (names.size() == 1) {
        const int rc = open(names[0], want_trace);
        if (rc == 0 && ranges && novl > 0) ranges->assign(1, std::make_pair(aread.front(), aread.back()));
        return rc;
    }
    std::vector<LasFile> parts(names.size());
    int64_t total = 0, ttotal = 0;
    for (size_t i = 0; i < names.size(); i++) {
        if (parts[i].open(names[i], want_trace) != 0) {
            error = parts[i].error;
            return -1;
        }
        if (i > 0 && parts[i].tspace != parts[0].tspace) {
            error = "parts of a split .las disagree on the trace spacing: " + names[i];
            return -1;
        }
        total += parts[i].novl;
        ttotal += parts[i].trace_off[(size_t)parts[i].novl];
    }
    novl = total;
    tspace = parts[0].tspace;
    tbytes = parts[0].tbytes;
    const size_t n = (size_t)total;
    aread.resize(n); bread.resize(n); abpos.resize(n); aepos.resize(n);
    bbpos.resize(n); bepos.resize(n); diffs.resize(n); flags.resize(n);
    trace_off.resize(n + 1);
    if (want_trace) trace.resize((size_t)ttotal);
    if (ranges) ranges->clear();
    size_t at = 0;
    int64_t tat = 0;
    for (LasFile& p : parts) {
        const size_t m = (size_t)p.novl;
        RawColumn<int32_t>* dst[8] = {&aread, &bread, &abpos, &aepos, &bbpos, &bepos, &diffs, &flags};
        RawColumn<int32_t>* src[8] = {&p.aread, &p.bread, &p.abpos, &p.aepos, &p.bbpos, &p.bepos, &p.diffs, &p.flags};
        for (int c = 0; c < 8; c++) memcpy(dst[c]->data() + at, src[c]->data(), m * sizeof(int32_t));
        for (size_t i = 0; i < m; i++) trace_off[at + i] = tat + p.trace_off[i];
        if (want_trace && p.trace_off[m] > 0) memcpy(trace.data() + tat, p.trace.data(), (size_t)p.trace_off[m]);
        if (ranges && m > 0) ranges->push_back(std::make_pair(p.aread.front(), p.aread.back()));
        at += m;
        tat += p.trace_off[m];
    }
    trace_off[n] = tat;
    return 0;
}

int LasFile::open(const std::string& las_name, bool want_trace) {
    int fd = ::open(las_name.c_str(), O_RDONLY);
    if (fd < 0) {
        error = "Cannot open " + las_name;
        return -1;
    }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 12) {
        ::close(fd);
        error = "short .las header in " + las_name;
        return -1;
    }
    const size_t fsize = (size_t)st.st_size;
    const uint8_t* base = (const uint8_t*)mmap(nullptr, fsize, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
    ::close(fd);
    if (base == MAP_FAILED) {
        error = "mmap failed for " + las_name;
        return -1;
    }
    memcpy(&novl, base, 8);
    memcpy(&tspace, base + 8, 4);
    tbytes = tspace <= 125 ? 1 : 2;
    auto fail = [&](const std::string& msg) {
        munmap((void*)base, fsize);
        error = msg;
        return -1;
    };
    if (novl < 0 || tspace <= 0) return fail("bad .las header in " + las_name);
    const bool timing = getenv("HINGE_B200_TIMING") != nullptr;
    struct timespec ts0;
    clock_gettime(CLOCK_MONOTONIC, &ts0);
    auto lap = [&](const char* what) {
        if (!timing) return;
        struct timespec t1;
        clock_gettime(CLOCK_MONOTONIC, &t1);
        fprintf(stderr, "[hinge_b200 timing]   las: %-22s %8.1f ms\n", what,
                1e3 * (double)(t1.tv_sec - ts0.tv_sec) + 1e-6 * (double)(t1.tv_nsec - ts0.tv_nsec));
        ts0 = t1;
    };

    // ---- the record walk, chunked when the file is big enough to pay for threads
    int T = (int)std::thread::hardware_concurrency();
    if (const char* v = getenv("HINGE_B200_IO_THREADS")) T = atoi(v);
    T = std::max(1, std::min(T, 32));
    size_t min_bytes = (size_t)8 << 20;  // below this the threads cost more than the walk
    if (const char* v = getenv("HINGE_B200_IO_MIN_BYTES")) min_bytes = (size_t)atoll(v);
    if (fsize < min_bytes) T = 1;
    std::vector<Walk> walks((size_t)T);
    bool joined = false;
    if (T > 1) {
        std::vector<size_t> start((size_t)T + 1, fsize);
        start[0] = 12;
        std::vector<std::thread> pool;
        for (int t = 1; t < T; t++)
            pool.emplace_back([&, t]() {
                const size_t lo = 12 + (fsize - 12) / T * t, hi = 12 + (fsize - 12) / T * (t + 1);
                start[t] = find_sync(base, fsize, lo, hi, tspace, tbytes);
            });
        for (auto& th : pool) th.join();
        pool.clear();
        // chunks without a sync point are absorbed by their left neighbour
        std::vector<size_t> stop((size_t)T, fsize);
        for (int t = T - 1; t >= 0; t--) {
            size_t s = fsize;
            for (int u = t + 1; u < T; u++)
                if (start[u] < fsize) {
                    s = start[u];
                    break;
                }
            stop[t] = s;
        }
        for (int t = 0; t < T; t++)
            pool.emplace_back([&, t]() {
                if (t > 0 && start[t] >= fsize) return;  // nothing of its own
                walk_records(base, fsize, start[t], stop[t], tbytes, &walks[t]);
            });
        for (auto& th : pool) th.join();
        joined = true;
        int64_t total = 0;
        for (int t = 0; t < T && joined; t++) {
            if (t > 0 && start[t] >= fsize) continue;
            const Walk& w = walks[t];
            joined = w.ok && w.end == stop[t];  // lands exactly on the neighbour's first record
            total += (int64_t)w.rec_off.size();
        }
        joined = joined && total == novl;
    }
    if (!joined) {  // one chase from the header (small files, or a guess that did not hold)
        T = 1;
        walks.assign(1, Walk());
        walks[0].rec_off.reserve((size_t)novl);
        walk_records(base, fsize, 12, fsize, tbytes, &walks[0]);
        // a well-formed file holds exactly novl records; trailing bytes are ignored like the reference does
        if (!walks[0].ok && (int64_t)walks[0].rec_off.size() < novl) return fail("truncated .las " + las_name);
        if ((int64_t)walks[0].rec_off.size() < novl) return fail("truncated .las " + las_name);
        walks[0].rec_off.resize((size_t)novl);
    }
    threads_used = T;
    lap(T > 1 ? "record walk (chunked)" : "record walk (one chase)");

    // ---- gather into the struct of arrays, every walk's records in parallel
    const size_t n = (size_t)novl;
    aread.resize(n); bread.resize(n); abpos.resize(n); aepos.resize(n);
    bbpos.resize(n); bepos.resize(n); diffs.resize(n); flags.resize(n);
    trace_off.resize(n + 1);
    std::vector<size_t> first((size_t)walks.size() + 1, 0);
    std::vector<int64_t> tfirst((size_t)walks.size() + 1, 0);
    for (size_t t = 0; t < walks.size(); t++) {
        first[t + 1] = first[t] + walks[t].rec_off.size();
        int64_t tb = 0;
        if (walks.size() == 1) {  // the sequential walk may have counted records past novl
            for (size_t i = 0; i < walks[0].rec_off.size(); i++) {
                int32_t tlen;
                memcpy(&tlen, base + walks[0].rec_off[i], 4);
                tb += (int64_t)tlen * tbytes;
            }
        } else {
            tb = walks[t].trace_bytes;
        }
        tfirst[t + 1] = tfirst[t] + tb;
    }
    const int64_t trace_total = tfirst[walks.size()];
    if (want_trace) trace.resize((size_t)trace_total);
    {
        // split every walk's records over a few workers so that T = 1 still fills in parallel
        int workers = (int)std::thread::hardware_concurrency();
        if (const char* v = getenv("HINGE_B200_IO_THREADS")) workers = atoi(v);
        workers = std::max(1, std::min(workers, 32));
        if (n < (1u << 16)) workers = 1;
        std::vector<std::thread> pool;
        auto fill = [&](size_t wi, size_t lo, size_t hi, int64_t tpos) {
            const Walk& w = walks[wi];
            for (size_t i = lo; i < hi; i++) {
                const uint8_t* r = base + w.rec_off[i];
                int32_t rec[9];
                memcpy(rec, r, 36);
                const size_t g = first[wi] + i;
                diffs[g] = rec[1]; abpos[g] = rec[2]; bbpos[g] = rec[3]; aepos[g] = rec[4];
                bepos[g] = rec[5]; flags[g] = rec[6]; aread[g] = rec[7]; bread[g] = rec[8];
                trace_off[g] = tpos;
                const size_t tb = (size_t)rec[0] * (size_t)tbytes;
                if (want_trace) memcpy(trace.data() + tpos, r + kRec, tb);
                tpos += (int64_t)tb;
            }
        };
        if (walks.size() > 1 || workers == 1) {
            for (size_t wi = 0; wi < walks.size(); wi++)
                pool.emplace_back(fill, wi, (size_t)0, walks[wi].rec_off.size(), tfirst[wi]);
        } else {
            // one walk, many workers: slice it; every slice needs its trace offset first
            const Walk& w = walks[0];
            const size_t per = (n + workers - 1) / workers;
            std::vector<int64_t> tpos((size_t)workers + 1, 0);
            for (int k = 0; k < workers; k++) {
                int64_t tb = 0;
                for (size_t i = std::min(n, per * k); i < std::min(n, per * (k + 1)); i++) {
                    int32_t tlen;
                    memcpy(&tlen, base + w.rec_off[i], 4);
                    tb += (int64_t)tlen * tbytes;
                }
                tpos[k + 1] = tpos[k] + tb;
            }
            for (int k = 0; k < workers; k++)
                pool.emplace_back(fill, (size_t)0, std::min(n, per * k), std::min(n, per * (k + 1)), tpos[k]);
        }
        for (auto& th : pool) th.join();
    }
    trace_off[n] = trace_total;
    lap("gather to columns");
    munmap((void*)base, fsize);
    lap("munmap");
    return 0;
}

// ---------------------------------------------------------------- INI

namespace {
std::string lower(std::string s) {
    std::transform(s.begin(), s.end(), s.begin(), ::tolower);
    return s;
}
std::string rstrip(std::string s) {
    while (!s.empty() && isspace((unsigned char)s.back())) s.pop_back();
    return s;
}
size_t lskip(const std::string& s, size_t i) {
    while (i < s.size() && isspace((unsigned char)s[i])) i++;
    return i;
}
// position of first `c`, or of a ';' that follows whitespace, or npos
size_t find_char_or_comment(const std::string& s, size_t i, char c) {
    bool ws = false;
    for (; i < s.size(); i++) {
        if (c && s[i] == c) return i;
        if (ws && s[i] == ';') return i;
        ws = isspace((unsigned char)s[i]) != 0;
    }
    return std::string::npos;
}
}  // namespace

int Ini::load(const std::string& path) {
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return -1;
    values_.clear();
    text_.clear();
    std::string section, prev_name;
    char linebuf[200];  // INI_MAX_LINE: longer lines are consumed in pieces
    int err = 0, lineno = 0;
    while (fgets(linebuf, sizeof linebuf, f)) {
        text_ += linebuf;
        lineno++;
        std::string raw = rstrip(linebuf);
        size_t b = lskip(raw, 0);
        std::string line = raw.substr(b);
        auto store = [&](const std::string& name, const std::string& value) {
            std::string key = lower(section + "=" + name);
            std::string& v = values_[key];
            if (!v.empty()) v += "\n";
            v += value;
        };
        if (line.empty()) continue;
        if (line[0] == ';' || line[0] == '#') continue;
        if (!prev_name.empty() && b > 0) {  // continuation of the previous value
            store(prev_name, line);
            continue;
        }
        if (line[0] == '[') {
            size_t e = find_char_or_comment(line, 1, ']');
            if (e != std::string::npos && line[e] == ']') {
                section = line.substr(1, e - 1).substr(0, 49);
                prev_name.clear();
            } else if (!err) {
                err = lineno;
            }
            continue;
        }
        size_t e = find_char_or_comment(line, 0, '=');
        if (e == std::string::npos || line[e] != '=') e = find_char_or_comment(line, 0, ':');
        if (e != std::string::npos && (line[e] == '=' || line[e] == ':')) {
            std::string name = rstrip(line.substr(0, e));
            size_t vb = lskip(line, e + 1);
            std::string value = line.substr(vb);
            size_t c = find_char_or_comment(value, 0, '\0');
            if (c != std::string::npos) value.resize(c);
            value = rstrip(value);
            prev_name = name.substr(0, 49);
            store(name, value);
        } else if (!err) {
            err = lineno;
        }
    }
    fclose(f);
    return err;
}

std::string Ini::get(const std::string& section, const std::string& name) const {
    auto it = values_.find(lower(section + "=" + name));
    return it == values_.end() ? std::string() : it->second;
}

long Ini::get_integer(const std::string& section, const std::string& name, long def) const {
    std::string v = get(section, name);
    char* end;
    long n = strtol(v.c_str(), &end, 0);
    return end > v.c_str() ? n : def;
}

double Ini::get_real(const std::string& section, const std::string& name, double def) const {
    std::string v = get(section, name);
    char* end;
    double n = strtod(v.c_str(), &end);
    return end > v.c_str() ? n : def;
}

bool Ini::get_boolean(const std::string& section, const std::string& name, bool def) const {
    std::string v = lower(get(section, name));
    if (v == "true" || v == "yes" || v == "on" || v == "1") return true;
    if (v == "false" || v == "no" || v == "off" || v == "0") return false;
    return def;
}

void load_filter_params(const Ini& ini, bool has_qv, hg_filter_params* p) {
    // filter.cpp:377-409
    p->min_cov = (int)ini.get_integer("filter", "min_cov", -1);
    p->cut_off = (int)ini.get_integer("filter", "cut_off", -1);
    p->theta = (int)ini.get_integer("filter", "theta", -1);
    p->est_cov = (int)ini.get_integer("filter", "ec", 0);
    p->reso = 40;
    p->use_qv_mask = ini.get_boolean("filter", "use_qv", true) && has_qv;
    p->use_coverage_mask = ini.get_boolean("filter", "coverage", true);
    p->coverage_fraction = (int)ini.get_integer("filter", "coverage_frac_repeat_annotation", 3);
    p->min_repeat_annotation_threshold =
        (int)ini.get_integer("filter", "min_repeat_annotation_threshold", 10);
    p->max_repeat_annotation_threshold =
        (int)ini.get_integer("filter", "max_repeat_annotation_threshold", 20);
    p->repeat_annotation_gap_threshold =
        (int)ini.get_integer("filter", "repeat_annotation_gap_threshold", 300);
    p->no_hinge_region = (int)ini.get_integer("filter", "no_hinge_region", 500);
    p->hinge_min_support = (int)ini.get_integer("filter", "hinge_min_support", 7);
    p->hinge_bin_pileup_threshold = (int)ini.get_integer("filter", "hinge_min_pileup", 7);
    p->hinge_read_unbridged_threshold = (int)ini.get_integer("filter", "hinge_unbridged", 6);
    p->hinge_tolerance_length = (int)ini.get_integer("filter", "hinge_tolerance_length", 100);
    p->hinge_bin_length = 2 * p->hinge_tolerance_length;  // filter.cpp:405 overrides hinge_bin
    p->delete_telomere = ini.get_integer("layout", "del_telomere", 0) != 0;
}

void load_layout_params(const Ini& ini, hg_layout_params* p) {
    // hinging.cpp:775-803 (maximal.cpp:443-474 reads the same [filter] keys)
    p->length_threshold = (int)ini.get_integer("filter", "length_threshold", -1);
    p->aln_threshold = (int)ini.get_integer("filter", "aln_threshold", -1);
    p->theta = (int)ini.get_integer("filter", "theta", -1);
    p->theta2 = (int)ini.get_integer("filter", "theta2", 0);
    p->use_two_matches = ini.get_integer("layout", "use_two_matches", 1) != 0;
    p->hinge_slack = (int)ini.get_integer("layout", "hinge_slack", 1000);
    p->hinge_tolerance = (int)ini.get_integer("layout", "hinge_tolerance", 150);
    p->kill_hinge_overlap = (int)ini.get_integer("layout", "kill_hinge_overlap", 300);
    p->kill_hinge_internal = (int)ini.get_integer("layout", "kill_hinge_internal", 40);
    p->matching_hinge_slack = (int)ini.get_integer("layout", "matching_hinge_slack", 200);
    p->num_events_telomere = (int)ini.get_integer("layout", "num_events_telomere", 7);
    p->min_connected_component_size =
        (int)ini.get_integer("layout", "min_connected_component_size", 8);
    p->keep_only_maximal =
        ini.get_integer("layout", "keep_only_matches_between_maximal_reads", 1) != 0;
    p->delete_telomeres = ini.get_integer("layout", "del_telomeres", 0) != 0;
}

// ---------------------------------------------------------------- TextOut

TextOut::TextOut(const std::string& path, bool append) : buf_(1 << 20) {
    fp_ = fopen(path.c_str(), append ? "a" : "w");
    if (!fp_) {  // an unwritable path must not end in fwrite(NULL): report it, swallow the output
        fprintf(stderr, "hinge_b200: cannot write %s\n", path.c_str());
        failed_ = true;
        fp_ = fopen("/dev/null", "w");
    }
}
TextOut::~TextOut() { close(); }

void TextOut::flush() {
    if (fp_ && len_) fwrite(buf_.data(), 1, len_, (FILE*)fp_);
    len_ = 0;
}

void TextOut::close() {
    if (fp_) {
        flush();
        fclose((FILE*)fp_);
        fp_ = nullptr;
    }
}

void TextOut::put_char(char c) {
    if (len_ + 1 > buf_.size()) flush();
    buf_[len_++] = c;
}

void TextOut::put_bytes(const char* s, size_t n) {
    flush();
    if (fp_ && n) fwrite(s, 1, n, (FILE*)fp_);
}

void TextOut::put_str(const char* s) {
    while (*s) put_char(*s++);
}

void TextOut::put_int(long v) {
    if (len_ + 24 > buf_.size()) flush();
    char tmp[24];
    int n = 0;
    unsigned long u = v < 0 ? 0ul - (unsigned long)v : (unsigned long)v;
    do {
        tmp[n++] = (char)('0' + u % 10);
        u /= 10;
    } while (u);
    if (v < 0) buf_[len_++] = '-';
    while (n) buf_[len_++] = tmp[--n];
}

}  // namespace hg
