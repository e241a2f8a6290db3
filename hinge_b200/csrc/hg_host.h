// Shared pieces of the file-level stage drivers (hg_host*.cpp).
#ifndef HG_HOST_H
#define HG_HOST_H
#include <string>

#include "../../include/hinge_b200.h"
#include "hg_io.h"

namespace hg {

// Flags of Reads_filter / get_maximal_reads / hinging
// (filter.cpp:172-183, maximal.cpp:242-253, hinging.cpp:621-640).
struct Args {
    std::string db, las, paf, config, fasta, prefix = "out", restrictreads, log = "log", out;
    bool mlas = false, debug = false;
};

bool parse_args(int argc, char** argv, bool layout, Args* a, std::string* err);
int check_inputs(const Args& a, std::string* las_name);
int load_inputs(const Args& a, bool want_trace, Ini* ini, ReadDB* db, LasFile* las);
int open_context(const ReadDB& db, const LasFile& las, bool with_trace, hg_ctx** ctx);

}  // namespace hg
#endif
