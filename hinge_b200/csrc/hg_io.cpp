// See hg_io.h.  Written from the on-disk formats, not from the reference's
// reader code.
#include "hg_io.h"

#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>

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
            if (fscanf(f, "  %9d %s %s\n", &last, a, b) != 3) {
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
    const uint8_t* base = (const uint8_t*)mmap(nullptr, fsize, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (base == MAP_FAILED) {
        error = "mmap failed for " + las_name;
        return -1;
    }
    madvise((void*)base, fsize, MADV_SEQUENTIAL);
    memcpy(&novl, base, 8);
    memcpy(&tspace, base + 8, 4);
    tbytes = tspace <= 125 ? 1 : 2;

    const size_t n = (size_t)novl;
    aread.resize(n); bread.resize(n); abpos.resize(n); aepos.resize(n);
    bbpos.resize(n); bepos.resize(n); diffs.resize(n); flags.resize(n);
    trace_off.assign(n + 1, 0);
    // record = 40 B: tlen diffs abpos bbpos aepos bepos flags aread bread pad
    // (align.h:126-132,332-337 minus the leading trace pointer, align.c:3042-3049)
    size_t p = 12;
    int64_t tpos = 0;
    for (size_t i = 0; i < n; i++) {
        if (p + 40 > fsize) {
            munmap((void*)base, fsize);
            error = "truncated .las " + las_name;
            return -1;
        }
        int32_t rec[9];
        memcpy(rec, base + p, 36);
        diffs[i] = rec[1]; abpos[i] = rec[2]; bbpos[i] = rec[3]; aepos[i] = rec[4];
        bepos[i] = rec[5]; flags[i] = rec[6]; aread[i] = rec[7]; bread[i] = rec[8];
        size_t tb = (size_t)rec[0] * (size_t)tbytes;
        p += 40;
        if (rec[0] < 0 || p + tb > fsize) {
            munmap((void*)base, fsize);
            error = "truncated .las " + las_name;
            return -1;
        }
        tpos += (int64_t)tb;
        trace_off[i + 1] = tpos;
        p += tb;
    }
    if (want_trace) {
        trace.resize((size_t)tpos);
        p = 12;
        for (size_t i = 0; i < n; i++) {
            size_t tb = (size_t)(trace_off[i + 1] - trace_off[i]);
            memcpy(trace.data() + trace_off[i], base + p + 40, tb);
            p += 40 + tb;
        }
    }
    munmap((void*)base, fsize);
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

TextOut::TextOut(const std::string& path) : buf_(1 << 20) { fp_ = fopen(path.c_str(), "w"); }
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
