// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See hinge_oracle.h.
// CLI with the reference's flag names so tests can run oracle, reference and
// product side by side:  hinge_oracle filter|maximal|layout --db X --las Y
//                        --config INI -x PREFIX [-o OUT]
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "hinge_oracle.h"

using namespace oracle;

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: hinge_oracle filter|maximal|layout --db X --las Y --config INI -x P [-o O]\n");
        return 1;
    }
    std::string cmd = argv[1], db, las, config, prefix = "out", outp;
    bool mlas = false;
    for (int i = 2; i < argc; i++) {
        std::string k = argv[i];
        if (k == "--mlas") {
            mlas = true;
            continue;
        }
        if (i + 1 >= argc) break;
        std::string v = argv[++i];
        if (k == "--db" || k == "-b") db = v;
        else if (k == "--las" || k == "-l") las = v;
        else if (k == "--config" || k == "-c") config = v;
        else if (k == "--prefix" || k == "-x") prefix = v;
        else if (k == "--out" || k == "-o") outp = v;
    }
    // --mlas: the parts <las>.1.las, <las>.2.las, ... (filter.cpp:35-63,228-235)
    std::vector<std::string> parts;
    if (mlas) {
        for (int i = 1;; i++) {
            std::string name = las + "." + std::to_string(i) + ".las";
            FILE* f = fopen(name.c_str(), "rb");
            if (!f) break;
            fclose(f);
            parts.push_back(name);
        }
        if (parts.empty()) {
            fprintf(stderr, "hinge_oracle: no parts %s.N.las\n", las.c_str());
            return 1;
        }
    } else {
        if (las.size() < 4 || las.substr(las.size() - 4) != ".las") las += ".las";
        parts.push_back(las);
    }
    Params p;
    Data d;
    std::string err;
    if (!load_ini(config, &p, &err) || !load_db(db, &d, &err)) {
        fprintf(stderr, "hinge_oracle: %s\n", err.c_str());
        return 1;
    }
    if (cmd == "filter") {
        // one pass per part, state carried like the reference does (hinge_oracle.h: FilterCarry)
        FilterCarry carry;
        for (size_t part = 0; part < parts.size(); part++) {
            if (!load_las(parts[part], &d, &err)) {
                fprintf(stderr, "hinge_oracle: %s\n", err.c_str());
                return 1;
            }
            if (d.novl == 0) {
                fprintf(stderr, "No alignments!\n");
                return 1;
            }
            FilterOut o;
            run_filter(d, p, &o, &carry);
            write_filter_files(o, p, d.n_read, prefix, mlas ? (int)part : -1);
        }
        return 0;
    }
    // maximal / layout: the parts hold disjoint, ascending A-read ranges and the reference's per-part
    // loops only share the per-read active flags, so the records are simply taken together
    std::vector<PII> ranges;
    for (size_t part = 0; part < parts.size(); part++) {
        const int64_t before = part ? d.novl : 0;
        if (!load_las(parts[part], &d, &err, part > 0)) {
            fprintf(stderr, "hinge_oracle: %s\n", err.c_str());
            return 1;
        }
        if (d.novl == before) {
            fprintf(stderr, "No alignments!\n");
            return 1;
        }
        ranges.push_back(PII(d.aread[before], d.aread[d.novl - 1]));
    }
    if (cmd == "maximal") {
        std::vector<PII> mask;
        read_mask_file(prefix + ".mas", d.n_read, &mask);
        MaximalOut o;
        run_maximal(d, p, mask, d.aread.front(), d.aread.back(), &o);
        write_maximal_files(o, ranges, prefix);
    } else if (cmd == "layout") {
        std::vector<PII> mask;
        std::vector<char> maximal;
        std::vector<std::vector<PII>> rep, hg;
        read_mask_file(prefix + ".mas", d.n_read, &mask);
        read_max_file(prefix + ".max", d.n_read, &maximal);
        read_pairs_file(prefix + ".repeat.txt", d.n_read, &rep);
        read_pairs_file(prefix + ".hinges.txt", d.n_read, &hg);
        LayoutOut o;
        run_layout(d, p, mask, maximal, rep, hg, &o);
        write_layout_files(o, d.n_read, prefix, outp);
    } else {
        fprintf(stderr, "unknown subcommand %s\n", cmd.c_str());
        return 1;
    }
    return 0;
}
