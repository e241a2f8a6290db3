// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// CPU restatement of the HINGE hot path (filter / maximal / layout) used as
// the parity oracle for the CUDA implementation.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may build, link or execute anything under oracle/.
//
// Language: C++ (not plain C) on purpose — the reference's results depend on
// the element order produced by libstdc++'s std::sort (introsort, unstable) and
// std::unordered_map iteration; the oracle calls those same library routines
// instead of restating them, so it inherits the reference's behaviour exactly
// when built with the same toolchain (g++ 13.3 here).
//
// Parity status: PINNED — tests/test_oracle_vs_reference.py checks the files
// this oracle writes byte-for-byte against the unmodified reference binaries
// (oracle/_ref, built by oracle/build_ref.sh) on daligner-made and synthetic
// fixtures, and tests/golden/ holds reference outputs for committed inputs.
#ifndef HINGE_ORACLE_H
#define HINGE_ORACLE_H
#include <stdint.h>

#include <string>
#include <utility>
#include <vector>

namespace oracle {

struct Params {
    // [filter] (filter.cpp:377-406)
    int length_threshold = -1, aln_threshold = -1, min_cov = -1, cut_off = -1, theta = -1;
    int theta2 = 0, est_cov = 0, reso = 40;
    bool use_qv = true, use_coverage = true;
    int coverage_fraction = 3, min_rep_thr = 10, max_rep_thr = 20, rep_gap = 300;
    int no_hinge_region = 500, hinge_min_support = 7, hinge_bin_pileup = 7;
    int hinge_unbridged = 6, hinge_tolerance_length = 100, hinge_bin_length = 200;
    bool del_telomere_filter = false;  // [layout] del_telomere  (filter.cpp:406)
    // [layout] (hinging.cpp:775-803)
    int hinge_slack = 1000, hinge_tolerance = 150, kill_hinge_overlap = 300;
    int kill_hinge_internal = 40, matching_hinge_slack = 200, num_events_telomere = 7;
    int min_cc_size = 8;
    bool use_two_matches = true, keep_only_maximal = true;
    bool del_telomeres_layout = false;  // [layout] del_telomeres (hinging.cpp:803)
    int threads = 1;  // host threads for the loops with independent iterations (results unchanged)
};

struct Data {
    int n_read = 0;
    std::vector<int> rlen;
    bool has_qv = false;
    std::vector<int64_t> qv_off;
    std::vector<uint8_t> qv;
    int64_t novl = 0;
    int tspace = 100, tbytes = 1;
    // raw .las columns (B coordinates still in DALIGNER convention)
    std::vector<int> aread, bread, abpos, aepos, bbpos, bepos, flags;
    std::vector<int64_t> trace_off;
    std::vector<uint8_t> trace;
};

typedef std::pair<int, int> PII;

struct FilterOut {
    int r_begin = 0, r_end = -1;
    int min_cov = 0, cov_est = 0;
    std::vector<std::vector<PII>> cov0, covc;  // (pos, coverage) per read
    std::vector<PII> mask, cmask;              // .mas / .cmas per read
    std::vector<std::vector<PII>> repeats, hinges;
    std::vector<int> cov_flag, self_flag;
};

struct MaximalOut {
    std::vector<char> active;  // after containment removal
    std::vector<PII> contained;  // (read, last containing read) in output order
};

struct Edge {
    int a, b, length, comp, type, weight;
    int eas, eae, ebs, ebe;  // trimmed match
    int ras, rae, rbs, rbe;  // effective read bounds
    int as, ae, bs, be;      // raw match (B flipped to forward strand)
    int hinge_pos;
};

struct LayoutOut {
    std::vector<int> garbage;
    std::vector<std::string> hgraph_lines;
    std::vector<std::vector<PII>> killed;  // (type,pos) order as printed
    std::vector<int> hinge_list;           // read,pos,type triples
    std::vector<Edge> edges;               // .edges.hinges order
    std::vector<Edge> skipped, greedy;
    std::vector<std::string> deadends;
};

bool load_ini(const std::string& path, Params* p, std::string* err);
bool load_db(const std::string& name, Data* d, std::string* err);
bool load_las(const std::string& name, Data* d, std::string* err, bool append = false);

// State the reference carries from one part of a multi-part run (--mlas, filter.cpp:474-1109) to the
// next: MIN_COV only ever grows (filter.cpp:677-678) and the masks of the reads of earlier parts stay
// known, while those of later parts are still (0,0) when a part calls its hinges (filter.cpp:534,884-889).
struct FilterCarry {
    int min_cov = 0;
    std::vector<PII> mask;
    bool started = false;
};
void run_filter(const Data& d, const Params& p, FilterOut* out, FilterCarry* carry = nullptr);
void run_maximal(const Data& d, const Params& p, const std::vector<PII>& mask, int r_begin,
                 int r_end, MaximalOut* out);
void run_layout(const Data& d, const Params& p, const std::vector<PII>& mask,
                const std::vector<char>& maximal, const std::vector<std::vector<PII>>& repeats,
                const std::vector<std::vector<PII>>& hinges, LayoutOut* out);

// text writers: byte-identical to the reference's files
// part < 0: single .las.  part >= 0: part number of a --mlas run: files are truncated by part 0 and
// appended to by the others; .repeat.txt is closed after part 0 (filter.cpp:1086) and every part's
// last read gets no .hinges.txt line (filter.cpp:1091).
void write_filter_files(const FilterOut& o, const Params& p, int n_read, const std::string& prefix, int part = -1);
// ranges: the [first, last] A-read of every part (one entry for a single .las): .max lists the
// surviving reads of those ranges only (maximal.cpp:873-878 runs inside the part loop)
void write_maximal_files(const MaximalOut& o, const std::vector<PII>& ranges, const std::string& prefix);
void write_layout_files(const LayoutOut& o, int n_read, const std::string& prefix,
                        const std::string& out_prefix);

// readers for the inter-stage files (same parsing rules as hinging.cpp:867-937)
void read_mask_file(const std::string& path, int n_read, std::vector<PII>* mask);
void read_max_file(const std::string& path, int n_read, std::vector<char>* maximal);
void read_pairs_file(const std::string& path, int n_read, std::vector<std::vector<PII>>* v);

}  // namespace oracle
#endif
