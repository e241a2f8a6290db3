// `hinge` front-end: the subcommands of the reference's dispatcher script that
// belong to the hot path (/root/reference/src/hinge:9-17).  Also answers to the
// reference's executable names, so the original `hinge` bash script can exec it
// through symlinks named Reads_filter / get_maximal_reads / hinging.
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <string>

#include "../../include/hinge_b200.h"

static int usage() {
    fprintf(stderr,
            "usage: hinge filter  --db DB --las LAS -x PREFIX --config INI\n"
            "       hinge maximal --db DB --las LAS -x PREFIX --config INI\n"
            "       hinge layout  --db DB --las LAS -x PREFIX --config INI -o OUT\n"
            "%s\n", hg_version());
    return 1;
}

// One stage per process: once its files are written nothing is left to do, so the process ends with
// _exit -- the CUDA context (gigabytes of allocations) goes back to the driver in one piece instead of
// being torn down call by call (hg_main_exit_after), and no static destructor runs.
static int finish(int rc) {
    fflush(NULL);
    _exit(rc);
}

int main(int argc, char** argv) {
    std::string self = argv[0];
    size_t slash = self.rfind('/');
    if (slash != std::string::npos) self = self.substr(slash + 1);
    hg_main_exit_after(1);
    if (self == "Reads_filter") return finish(hg_main_filter(argc, argv));
    if (self == "get_maximal_reads") return finish(hg_main_maximal(argc, argv));
    if (self == "hinging") return finish(hg_main_layout(argc, argv));
    if (argc < 2) return usage();
    std::string cmd = argv[1];
    argv[1] = argv[0];
    if (cmd == "filter") return finish(hg_main_filter(argc - 1, argv + 1));
    if (cmd == "maximal") return finish(hg_main_maximal(argc - 1, argv + 1));
    if (cmd == "layout") return finish(hg_main_layout(argc - 1, argv + 1));
    fprintf(stderr, "hinge: subcommand '%s' is not part of the B200 hot path (see DESIGN.md)\n", cmd.c_str());
    return usage();
}
