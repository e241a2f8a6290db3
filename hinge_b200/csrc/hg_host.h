// Shared pieces of the file-level stage drivers (hg_host*.cpp).
#ifndef HG_HOST_H
#define HG_HOST_H
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/hinge_b200.h"
#include "hg_io.h"

namespace hg {

// Flags of Reads_filter / get_maximal_reads / hinging
// (filter.cpp:172-183, maximal.cpp:242-253, hinging.cpp:621-640).
struct Args {
    std::string db, las, paf, config, fasta, prefix = "out", restrictreads, log = "log", out;
    bool mlas = false, debug = false;
};

// HINGE_B200_TIMING=1: wall time of the phases of a stage driver on stderr
struct PhaseTimer {
    bool on = getenv("HINGE_B200_TIMING") != nullptr;
    struct timespec t0;
    PhaseTimer() { clock_gettime(CLOCK_MONOTONIC, &t0); }
    void lap(const char* what) {
        if (!on) return;
        struct timespec t1;
        clock_gettime(CLOCK_MONOTONIC, &t1);
        fprintf(stderr, "[hinge_b200 timing] %-28s %8.1f ms\n", what,
                1e3 * (double)(t1.tv_sec - t0.tv_sec) + 1e-6 * (double)(t1.tv_nsec - t0.tv_nsec));
        t0 = t1;
    }
    // a quantity next to the phase times, same line format (bench.py parses both)
    void note(const char* what, double value, const char* unit) {
        if (on) fprintf(stderr, "[hinge_b200 timing] %-28s %8.1f %s\n", what, value, unit);
    }
};

bool parse_args(int argc, char** argv, bool layout, Args* a, std::string* err);
// las_names: the single <las>[.las], or the parts <las>.1.las, <las>.2.las, ... of a --mlas run
// (filter.cpp:35-63,228-241)
int check_inputs(const Args& a, std::vector<std::string>* las_names);
// Reads the DB, the INI and (load_las) all records -- the parts of a --mlas run taken together, with
// part_ranges = [first, last] A-read of every part.  Starts the creation of the CUDA context beside it.
int load_inputs(const Args& a, bool want_trace, Ini* ini, ReadDB* db, LasFile* las,
                std::vector<std::pair<int32_t, int32_t>>* part_ranges = nullptr, bool load_las = true);
int open_context(const ReadDB& db, const LasFile& las, bool with_trace, hg_ctx** ctx);
// load_inputs starts the creation of the CUDA context on a side thread; every error return between
// load_inputs and open_context must give it back (joins the thread, destroys the context)
void drop_early_context();
// hg_ctx_destroy, unless the process is about to exit anyway (hg_main_exit_after)
void release_context(hg_ctx* ctx);

}  // namespace hg
#endif
