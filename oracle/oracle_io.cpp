// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See hinge_oracle.h.
// Minimal loaders for the oracle: plain fread over the documented layouts
// (include/DB.h:214-303, include/align.h:126-132,332-337) and a small INI
// reader with inih's comment rule (lib/ini.c:45-54).
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>

#include "hinge_oracle.h"

namespace oracle {

static std::string strip(const std::string& s) {
    size_t b = 0, e = s.size();
    while (b < e && isspace((unsigned char)s[b])) b++;
    while (e > b && isspace((unsigned char)s[e - 1])) e--;
    return s.substr(b, e - b);
}

bool load_ini(const std::string& path, Params* p, std::string* err) {
    FILE* f = fopen(path.c_str(), "r");
    if (!f) {
        *err = "Can't load " + path;
        return false;
    }
    std::map<std::string, std::string> kv;
    char buf[4096];
    std::string section;
    while (fgets(buf, sizeof buf, f)) {
        std::string line = strip(buf);
        if (line.empty() || line[0] == ';' || line[0] == '#') continue;
        if (line[0] == '[') {
            size_t e = line.find(']');
            if (e != std::string::npos) section = line.substr(1, e - 1);
            continue;
        }
        size_t eq = line.find_first_of("=:");
        if (eq == std::string::npos) continue;
        std::string name = strip(line.substr(0, eq)), value = strip(line.substr(eq + 1));
        for (size_t i = 1; i < value.size(); i++)  // ';' only after whitespace is a comment
            if (value[i] == ';' && isspace((unsigned char)value[i - 1])) {
                value = strip(value.substr(0, i));
                break;
            }
        std::string key = section + "=" + name;
        std::transform(key.begin(), key.end(), key.begin(), ::tolower);
        // lib/INIReader.cpp:74-81: a repeated key APPENDS "\n" + value, so numeric getters see the first one
        if (kv.count(key) && !kv[key].empty())
            kv[key] += "\n" + value;
        else
            kv[key] = value;
    }
    fclose(f);
    auto geti = [&](const char* k, int def) {
        auto it = kv.find(k);
        if (it == kv.end()) return def;
        char* end;
        long v = strtol(it->second.c_str(), &end, 0);
        return end > it->second.c_str() ? (int)v : def;
    };
    auto getb = [&](const char* k, bool def) {
        auto it = kv.find(k);
        if (it == kv.end()) return def;
        std::string v = it->second;
        std::transform(v.begin(), v.end(), v.begin(), ::tolower);
        if (v == "true" || v == "yes" || v == "on" || v == "1") return true;
        if (v == "false" || v == "no" || v == "off" || v == "0") return false;
        return def;
    };
    p->length_threshold = geti("filter=length_threshold", -1);
    p->aln_threshold = geti("filter=aln_threshold", -1);
    p->min_cov = geti("filter=min_cov", -1);
    p->cut_off = geti("filter=cut_off", -1);
    p->theta = geti("filter=theta", -1);
    p->theta2 = geti("filter=theta2", 0);
    p->est_cov = geti("filter=ec", 0);
    p->use_qv = getb("filter=use_qv", true);
    p->use_coverage = getb("filter=coverage", true);
    p->coverage_fraction = geti("filter=coverage_frac_repeat_annotation", 3);
    p->min_rep_thr = geti("filter=min_repeat_annotation_threshold", 10);
    p->max_rep_thr = geti("filter=max_repeat_annotation_threshold", 20);
    p->rep_gap = geti("filter=repeat_annotation_gap_threshold", 300);
    p->no_hinge_region = geti("filter=no_hinge_region", 500);
    p->hinge_min_support = geti("filter=hinge_min_support", 7);
    p->hinge_bin_pileup = geti("filter=hinge_min_pileup", 7);
    p->hinge_unbridged = geti("filter=hinge_unbridged", 6);
    p->hinge_tolerance_length = geti("filter=hinge_tolerance_length", 100);
    p->hinge_bin_length = 2 * p->hinge_tolerance_length;
    p->del_telomere_filter = geti("layout=del_telomere", 0) != 0;
    p->hinge_slack = geti("layout=hinge_slack", 1000);
    p->hinge_tolerance = geti("layout=hinge_tolerance", 150);
    p->kill_hinge_overlap = geti("layout=kill_hinge_overlap", 300);
    p->kill_hinge_internal = geti("layout=kill_hinge_internal", 40);
    p->matching_hinge_slack = geti("layout=matching_hinge_slack", 200);
    p->num_events_telomere = geti("layout=num_events_telomere", 7);
    p->min_cc_size = geti("layout=min_connected_component_size", 8);
    p->use_two_matches = geti("layout=use_two_matches", 1) != 0;
    p->keep_only_maximal = geti("layout=keep_only_matches_between_maximal_reads", 1) != 0;
    p->del_telomeres_layout = geti("layout=del_telomeres", 0) != 0;
    return true;
}

static bool slurp(const std::string& path, std::vector<uint8_t>* out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseeko(f, 0, SEEK_END);
    size_t n = (size_t)ftello(f);
    fseeko(f, 0, SEEK_SET);
    out->resize(n);
    bool ok = n == 0 || fread(out->data(), 1, n, f) == n;
    fclose(f);
    return ok;
}

bool load_db(const std::string& name, Data* d, std::string* err) {
    std::string s = name;
    if (s.size() > 3 && s.substr(s.size() - 3) == ".db") s.resize(s.size() - 3);
    size_t slash = s.rfind('/');
    std::string dir = slash == std::string::npos ? "." : s.substr(0, slash);
    std::string root = slash == std::string::npos ? s : s.substr(slash + 1);
    int cutoff = 0, all = 1;
    {
        FILE* f = fopen((dir + "/" + root + ".db").c_str(), "r");
        if (!f) {
            *err = "cannot open " + name + ".db";
            return false;
        }
        char line[20000];
        while (fgets(line, sizeof line, f)) {
            long long size;
            if (sscanf(line, "size = %lld cutoff = %d all = %d", &size, &cutoff, &all) == 3) break;
        }
        fclose(f);
    }
    std::vector<uint8_t> idx;
    if (!slurp(dir + "/." + root + ".idx", &idx) || idx.size() < 112) {
        *err = "bad .idx";
        return false;
    }
    int ureads = *(int*)&idx[0], treads = *(int*)&idx[4];
    bool trimmed = !(cutoff <= 0 && all);
    int allflag = all ? 0 : 0x800;
    std::vector<int> kept;
    d->rlen.clear();
    for (int i = 0; i < ureads; i++) {
        const uint8_t* r = &idx[112 + 40 * (size_t)i];
        int rlen = *(const int*)(r + 4), flags = *(const int*)(r + 32);
        if (!trimmed || ((flags & 0x800) >= allflag && rlen >= cutoff)) {
            kept.push_back(i);
            d->rlen.push_back(rlen);
        }
    }
    d->n_read = (int)d->rlen.size();
    d->has_qv = false;
    std::vector<uint8_t> anno, data;
    if (slurp(dir + "/." + root + ".qual.anno", &anno) && anno.size() >= 8 &&
        slurp(dir + "/." + root + ".qual.data", &data)) {
        int tracklen = *(int*)&anno[0], size = *(int*)&anno[4];
        if (size == 0) size = 8;
        bool untrimmed_track = tracklen == ureads;
        if (size == 8 && (untrimmed_track || (tracklen == treads && tracklen == d->n_read))) {
            const int64_t* off = (const int64_t*)&anno[8];
            d->qv_off.assign(d->n_read + 1, 0);
            d->qv.clear();
            for (int j = 0; j < d->n_read; j++) {
                int src = untrimmed_track ? kept[j] : j;
                d->qv.insert(d->qv.end(), data.begin() + off[src], data.begin() + off[src + 1]);
                d->qv_off[j + 1] = (int64_t)d->qv.size();
            }
            d->has_qv = true;
        }
    }
    return true;
}

bool load_las(const std::string& name, Data* d, std::string* err, bool append) {
    FILE* f = fopen(name.c_str(), "rb");
    if (!f) {
        *err = "cannot open " + name;
        return false;
    }
    int64_t novl;
    int tspace;
    if (fread(&novl, 8, 1, f) != 1 || fread(&tspace, 4, 1, f) != 1) {
        fclose(f);
        *err = "short .las";
        return false;
    }
    const int64_t base = append ? d->novl : 0;  // parts of a split .las: appended in order
    if (!append) {
        d->trace_off.assign(1, 0);
        d->trace.clear();
    }
    novl += base;
    d->novl = novl;
    d->tspace = tspace;
    d->tbytes = tspace <= 125 ? 1 : 2;
    d->aread.resize(novl); d->bread.resize(novl); d->abpos.resize(novl); d->aepos.resize(novl);
    d->bbpos.resize(novl); d->bepos.resize(novl); d->flags.resize(novl);
    d->trace_off.resize(novl + 1, 0);
    for (int64_t k = base; k < novl; k++) {
        int rec[10];
        if (fread(rec, 40, 1, f) != 1) {
            fclose(f);
            *err = "truncated .las";
            return false;
        }
        d->abpos[k] = rec[2]; d->bbpos[k] = rec[3]; d->aepos[k] = rec[4]; d->bepos[k] = rec[5];
        d->flags[k] = rec[6]; d->aread[k] = rec[7]; d->bread[k] = rec[8];
        size_t tb = (size_t)rec[0] * d->tbytes, old = d->trace.size();
        d->trace.resize(old + tb);
        if (tb && fread(&d->trace[old], tb, 1, f) != 1) {
            fclose(f);
            *err = "truncated .las";
            return false;
        }
        d->trace_off[k + 1] = (int64_t)d->trace.size();
    }
    fclose(f);
    return true;
}

}  // namespace oracle
