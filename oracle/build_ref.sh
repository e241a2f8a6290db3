#!/usr/bin/env bash
# Builds the UNMODIFIED reference stage binaries (Reads_filter,
# get_maximal_reads, hinging) and the upstream fixture tools (simulator,
# fasta2DB, DBsplit, daligner, LAsort, LAmerge, LAcheck, DASqv) from the
# sources where they lie under /root/reference, into oracle/_ref/bin.
# Test infrastructure only: nothing under oracle/ is linked into the product.
#
# The reference's own CMake build is not used (needs Boost, which is absent);
# the three stages compile from their own few source files.  hinging.cpp needs
# <boost/graph/...>: oracle/boost_shim provides the four calls it uses.
# Flags follow the reference's effective Release build
# (src/CMakeLists.txt:22-24, src/spdlog/CMakeLists.txt:14-15).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${HINGE_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
BIN="$OUT/bin"
OBJ="$OUT/obj"
if [ ! -d "$REF/src" ]; then
    echo "build_ref.sh: $REF not present; keeping prebuilt oracle/_ref" >&2
    exit 0
fi
mkdir -p "$BIN" "$OBJ"
R="$REF/src"
CF="-O3 -DNDEBUG -fopenmp -w"
for f in DB QV align ini paf; do
    gcc $CF -I"$R/include" -c "$R/lib/$f.c" -o "$OBJ/$f.o"
done
for f in LAInterface INIReader; do
    g++ $CF -std=gnu++11 -I"$R/include" -I"$R/spdlog/include" -c "$R/lib/$f.cpp" -o "$OBJ/$f.o"
done
LIBO="$OBJ/DB.o $OBJ/QV.o $OBJ/align.o $OBJ/ini.o $OBJ/paf.o $OBJ/LAInterface.o $OBJ/INIReader.o"
build_stage() {  # src exe
    g++ $CF -std=gnu++11 -I"$HERE/boost_shim" -I"$R/include" -I"$R/spdlog/include" \
        "$R/$1" $LIBO -lz -lpthread -o "$BIN/$2"
}
build_stage filter/filter.cpp Reads_filter &
build_stage maximal/maximal.cpp get_maximal_reads &
build_stage layout/hinging.cpp hinging &
wait

# upstream tools: their Makefiles build in place, so work on a scratch copy
if [ "${HINGE_REF_TOOLS:-1}" = "1" ]; then
    TP="$(mktemp -d /tmp/hinge_tp.XXXXXX)"
    cp -r "$REF/thirdparty/DAZZ_DB" "$REF/thirdparty/DALIGNER" "$REF/thirdparty/DASCRUBBER" "$TP/"
    chmod -R u+w "$TP"
    make -s -C "$TP/DAZZ_DB" simulator fasta2DB DBsplit DBshow DBdump >/dev/null 2>&1 || true
    make -s -C "$TP/DALIGNER" daligner LAsort LAmerge LAsplit LAcheck LAshow LAdump >/dev/null 2>&1 || true
    make -s -C "$TP/DASCRUBBER" DASqv >/dev/null 2>&1 || true
    for t in DAZZ_DB/simulator DAZZ_DB/fasta2DB DAZZ_DB/DBsplit DAZZ_DB/DBshow DAZZ_DB/DBdump \
             DALIGNER/daligner DALIGNER/LAsort DALIGNER/LAmerge DALIGNER/LAsplit DALIGNER/LAcheck \
             DALIGNER/LAshow DALIGNER/LAdump DASCRUBBER/DASqv; do
        [ -x "$TP/$t" ] && cp "$TP/$t" "$BIN/" || echo "build_ref.sh: tool $t not built" >&2
    done
    rm -rf "$TP"
fi
cp "$REF/utils/nominal.ini" "$OUT/nominal.ini"
ls "$BIN"
