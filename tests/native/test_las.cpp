// The chunked .las ingest (several walkers that guess record starts and are then checked to
// join up) must return exactly what one sequential walk returns, whatever the thread count,
// with and without traces, and must fall back cleanly when a guess cannot be made.
//   usage: test_las <file.las> [<file.las> ...]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "../../hinge_b200/csrc/hg_io.h"

static bool same(const hg::LasFile& a, const hg::LasFile& b, bool trace) {
    if (a.novl != b.novl || a.tspace != b.tspace || a.tbytes != b.tbytes) return false;
    const size_t n = (size_t)a.novl;
#define COL(c) if (memcmp(a.c.data(), b.c.data(), sizeof(a.c[0]) * n)) return false;
    COL(aread) COL(bread) COL(abpos) COL(aepos) COL(bbpos) COL(bepos) COL(diffs) COL(flags)
#undef COL
    if (memcmp(a.trace_off.data(), b.trace_off.data(), 8 * (n + 1))) return false;
    if (trace && memcmp(a.trace.data(), b.trace.data(), (size_t)a.trace_off[n])) return false;
    return true;
}

int main(int argc, char** argv) {
    int checked = 0, chunked = 0;
    for (int f = 1; f < argc; f++) {
        for (int trace = 0; trace < 2; trace++) {
            setenv("HINGE_B200_IO_THREADS", "1", 1);
            hg::LasFile ref;
            if (ref.open(argv[f], trace != 0) != 0) {
                printf("FAIL open %s: %s\n", argv[f], ref.error.c_str());
                return 1;
            }
            for (const char* t : {"2", "3", "8", "13"}) {
                setenv("HINGE_B200_IO_THREADS", t, 1);
                setenv("HINGE_B200_IO_MIN_BYTES", "0", 1);  // chunk even small files
                hg::LasFile got;
                if (got.open(argv[f], trace != 0) != 0 || !same(ref, got, trace != 0)) {
                    printf("FAIL %s threads=%s trace=%d walkers=%d\n", argv[f], t, trace, got.threads_used);
                    return 1;
                }
                checked++;
                chunked += got.threads_used > 1;
            }
        }
    }
    if (chunked == 0) {
        printf("FAIL: no run was actually chunked\n");
        return 1;
    }
    printf("OK %d comparisons, %d chunked\n", checked, chunked);
    return 0;
}
