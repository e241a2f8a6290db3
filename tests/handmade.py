"""Hand-made DAZZ_DB + .las fixtures for edge cases the generators do not reach
(many overlaps per read pair, 16-bit traces, reads without overlaps, self-overlaps).
Formats: SURVEY.md Appendix B (DB.h:214-303, align.h:126-132,332-337)."""
import os
import struct

import numpy as np


def make_trace(abpos, aepos, bbpos, bepos, tspace, rng):
    """(diff, bdelta) per tspace segment of A; b-deltas add up to the B span."""
    nseg = (aepos - 1) // tspace - abpos // tspace + 1
    bounds = [abpos] + [(abpos // tspace + 1 + j) * tspace for j in range(nseg - 1)] + [aepos]
    blen, alen = bepos - bbpos, aepos - abpos
    out, acc = [], 0
    for j in range(nseg):
        tgt = (bounds[j + 1] - abpos) * blen // max(alen, 1)
        bd = blen - acc if j == nseg - 1 else tgt - acc
        if 0 < j < nseg - 2:
            bd += int(rng.integers(-2, 3))
        bd = max(bd, 0)
        if j == nseg - 1:
            bd = blen - acc
        acc += bd
        out += [int(rng.integers(0, 20)), bd]
    assert acc == blen and min(out) >= 0
    return out


def write_fixture(directory, root, rlen, records, tspace=100, qv=None, seed=0):
    """records: list of (aread, bread, abpos, aepos, bbpos, bepos, comp) with B coordinates in
    DALIGNER convention (complement strand when comp); sorted here like LAsort does."""
    rng = np.random.default_rng(seed)
    n = len(rlen)
    os.makedirs(directory, exist_ok=True)
    with open(os.path.join(directory, root + ".db"), "w") as f:
        f.write("files = %9d\n" % 1)
        f.write("  %9d %s %s\n" % (n, root, "Hand"))
        f.write("blocks = %9d\n" % 1)
        f.write("size = %9d cutoff = %9d all = %1d\n" % (400, 0, 1))
        f.write(" %9d %9d\n" % (0, 0))
        f.write(" %9d %9d\n" % (n, n))
    hdr = bytearray(112)
    struct.pack_into("<4i", hdr, 0, n, n, 0, 1)
    struct.pack_into("<4f", hdr, 16, .25, .25, .25, .25)
    struct.pack_into("<i", hdr, 32, max(rlen))
    struct.pack_into("<q", hdr, 40, sum(rlen))
    boff = 0
    with open(os.path.join(directory, "." + root + ".idx"), "wb") as f:
        f.write(hdr)
        for i, rl in enumerate(rlen):
            r = bytearray(40)
            struct.pack_into("<3i", r, 0, i + 1, rl, 0)
            struct.pack_into("<2q", r, 16, boff, -1)
            struct.pack_into("<i", r, 32, 0x800 | 850)
            f.write(r)
            boff += (rl + 3) >> 2
    with open(os.path.join(directory, "." + root + ".bps"), "wb") as f:
        f.truncate(boff)
    if qv is not None:
        off = np.zeros(n + 1, np.int64)
        for i, rl in enumerate(rlen):
            off[i + 1] = off[i] + (rl + tspace - 1) // tspace
        data = np.full(off[-1], 20, np.uint8) if qv == "good" else np.asarray(qv, np.uint8)
        with open(os.path.join(directory, "." + root + ".qual.anno"), "wb") as f:
            f.write(struct.pack("<2i", n, 8))
            f.write(off.tobytes())
        with open(os.path.join(directory, "." + root + ".qual.data"), "wb") as f:
            f.write(data.tobytes())
    records = sorted(records, key=lambda r: (r[0], r[1], r[2]))
    tb = 1 if tspace <= 125 else 2
    with open(os.path.join(directory, root + ".las"), "wb") as f:
        f.write(struct.pack("<qi", len(records), tspace))
        for a, b, ab, ae, bb, be, comp in records:
            assert 0 <= ab < ae <= rlen[a] and 0 <= bb < be <= rlen[b]
            tr = make_trace(ab, ae, bb, be, tspace, rng)
            f.write(struct.pack("<10i", len(tr), sum(tr[0::2]), ab, bb, ae, be, comp, a, b, 0))
            f.write(np.asarray(tr, np.uint8 if tb == 1 else np.uint16).tobytes())
    return len(records)


def both_directions(a, b, ab, ae, bb, be, comp, rlen):
    """The record and its mirror (B as the A-read), like daligner emits."""
    if comp:
        # A interval complemented in the mirror record
        return [(a, b, ab, ae, bb, be, 1), (b, a, rlen[b] - be, rlen[b] - bb, rlen[a] - ae, rlen[a] - ab, 1)]
    return [(a, b, ab, ae, bb, be, 0), (b, a, bb, be, ab, ae, 0)]
