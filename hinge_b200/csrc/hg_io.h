// Host-side readers for the file formats at the drop-in boundary:
//   DAZZ_DB stub/index/track files, DALIGNER .las files, the INI config.
// Formats follow /root/reference/src/include/DB.h:214-303,
// /root/reference/src/include/align.h:126-132,332-337 and
// /root/reference/src/lib/ini.c; the code is written from the format, it
// shares nothing with the reference's readers.
#ifndef HG_IO_H
#define HG_IO_H
#include <stdint.h>
#include <stdlib.h>
#include <map>
#include <string>
#include <vector>

#include "hg_params.h"

namespace hg {

// Reads of a (trimmed) DAZZ_DB, indexed the way DALIGNER numbers them.
struct ReadDB {
    int32_t n_read = 0;
    std::vector<int32_t> rlen;
    bool has_qv = false;          // `qual` track present and in sync
    std::vector<int64_t> qv_off;  // n_read + 1 offsets into qv
    std::vector<uint8_t> qv;      // one intrinsic QV per tspace tile
    std::string error;

    // Mirrors LAInterface::openDB + getReadNumber + getQV
    // (/root/reference/src/lib/LAInterface.cpp:133-185,2556-2558,4369-4494).
    // Returns 0 on success; on failure `error` says why (the reference exits 1).
    int open(const std::string& db_name);
};

// Column buffer without value-initialisation (a std::vector would zero 28 B per record first,
// single-threaded, before the ingest overwrites them).
template <class T>
struct RawColumn {
    T* p = nullptr;
    size_t n = 0;
    RawColumn() {}
    RawColumn(const RawColumn&) = delete;
    RawColumn& operator=(const RawColumn&) = delete;
    ~RawColumn() { free(p); }
    void resize(size_t count) {  // contents undefined
        free(p);
        p = static_cast<T*>(malloc((count ? count : 1) * sizeof(T)));
        n = count;
    }
    T* data() { return p; }
    const T* data() const { return p; }
    size_t size() const { return n; }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
    const T& front() const { return p[0]; }
    const T& back() const { return p[n - 1]; }
};

// A whole .las file as a struct of arrays (the layout the kernels consume).
struct LasFile {
    int64_t novl = 0;
    int32_t tspace = 0;
    int32_t tbytes = 1;  // 1 if tspace <= 125 (align.h:58 TRACE_XOVR), else 2
    RawColumn<int32_t> aread, bread, abpos, aepos, bbpos, bepos, diffs, flags;
    RawColumn<int64_t> trace_off;  // novl + 1 byte offsets into trace
    RawColumn<uint8_t> trace;      // raw trace bytes, (diff, bdelta) pairs
    std::string error;
    int threads_used = 1;          // how the record walk went: > 1 = chunked, verified

    // Mirrors LAInterface::openAlignmentFile + getOverlap(0, n_read)
    // (/root/reference/src/lib/LAInterface.cpp:595-621,1519-1634).
    int open(const std::string& las_name, bool want_trace);
    // The parts of a split .las (--mlas: <base>.1.las, <base>.2.las, ...) taken together, in order;
    // ranges gets the [first, last] A-read of every part.
    int open_parts(const std::vector<std::string>& names, bool want_trace,
                   std::vector<std::pair<int32_t, int32_t>>* ranges);
};

// inih-compatible INI reader (/root/reference/src/lib/ini.c, INIReader.cpp):
// ';' starts a comment only after whitespace, so "1000;" keeps its ';' and
// strtol stops there, while "true;" fails the boolean match and falls back to
// the default.
class Ini {
public:
    // <0: file could not be opened (the only case the reference treats as fatal)
    int load(const std::string& path);
    long get_integer(const std::string& section, const std::string& name, long def) const;
    double get_real(const std::string& section, const std::string& name, double def) const;
    bool get_boolean(const std::string& section, const std::string& name, bool def) const;
    const std::string& text() const { return text_; }

private:
    std::string get(const std::string& section, const std::string& name) const;
    std::map<std::string, std::string> values_;
    std::string text_;
};

void load_filter_params(const Ini& ini, bool has_qv, hg_filter_params* p);
void load_layout_params(const Ini& ini, hg_layout_params* p);

// Buffered text emitter for the line-oriented outputs (integers separated by
// single characters); much faster than iostream, same bytes.
class TextOut {
public:
    explicit TextOut(const std::string& path, bool append = false);
    ~TextOut();
    bool ok() const { return fp_ != nullptr && !failed_; }
    void put_int(long v);
    void put_char(char c);
    void put_str(const char* s);
    void put_bytes(const char* s, size_t n);  // flushes, then writes straight through
    void close();

private:
    void flush();
    void* fp_ = nullptr;
    bool failed_ = false;
    std::vector<char> buf_;
    size_t len_ = 0;
};

}  // namespace hg
#endif
