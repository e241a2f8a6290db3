// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See hinge_oracle.h.
// CLI with the reference's flag names so tests can run oracle, reference and
// product side by side:  hinge_oracle filter|maximal|layout --db X --las Y
//                        --config INI -x PREFIX [-o OUT]
#include <stdio.h>
#include <string.h>

#include <string>

#include "hinge_oracle.h"

using namespace oracle;

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: hinge_oracle filter|maximal|layout --db X --las Y --config INI -x P [-o O]\n");
        return 1;
    }
    std::string cmd = argv[1], db, las, config, prefix = "out", outp;
    for (int i = 2; i + 1 < argc; i += 2) {
        std::string k = argv[i], v = argv[i + 1];
        if (k == "--db" || k == "-b") db = v;
        else if (k == "--las" || k == "-l") las = v;
        else if (k == "--config" || k == "-c") config = v;
        else if (k == "--prefix" || k == "-x") prefix = v;
        else if (k == "--out" || k == "-o") outp = v;
    }
    if (las.size() < 4 || las.substr(las.size() - 4) != ".las") las += ".las";
    Params p;
    Data d;
    std::string err;
    if (!load_ini(config, &p, &err) || !load_db(db, &d, &err) || !load_las(las, &d, &err)) {
        fprintf(stderr, "hinge_oracle: %s\n", err.c_str());
        return 1;
    }
    if (d.novl == 0) {
        fprintf(stderr, "No alignments!\n");
        return 1;
    }
    if (cmd == "filter") {
        FilterOut o;
        run_filter(d, p, &o);
        write_filter_files(o, p, d.n_read, prefix);
    } else if (cmd == "maximal") {
        std::vector<PII> mask;
        read_mask_file(prefix + ".mas", d.n_read, &mask);
        MaximalOut o;
        run_maximal(d, p, mask, d.aread.front(), d.aread.back(), &o);
        write_maximal_files(o, d.aread.front(), d.aread.back(), prefix);
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
