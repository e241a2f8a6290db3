"""Shared helpers of the test-suite: native builds, fixture materialisation,
running the three implementations (product CLI, oracle CLI, reference binaries)
and byte-comparing their output files."""
import hashlib
import json
import lzma
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")
INI = os.path.join(GOLDEN, "nominal.ini")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
ORACLE = os.path.join(ROOT, "oracle", "_build", "hinge_oracle")
SYNTH = os.path.join(ROOT, "tools", "_build", "hinge_synth")
HINGE = os.path.join(ROOT, "hinge_b200", "_build", "hinge")
LIB = os.path.join(ROOT, "hinge_b200", "_build", "libhinge_b200.so")

FILTER_OUT = ["mas", "cmas", "coverage.txt", "repeat.txt", "hinges.txt", "cov.flag", "self.flag"]
MAXIMAL_OUT = ["max", "contained.txt"]
LAYOUT_OUT = ["edges.hinges", "edges.hinges2", "hinge.list", "hgraph", "killed.hinges", "edges.skipped",
              "edges.greedy", "edges.1", "edges.2", "deadends.txt", "garbage.txt"]
FIXTURES = ["dal_small", "synth_small", "synth_long", "synth_noqv", "synth_frag"]


def build_all():
    import __graft_entry__ as ge

    ge.build_native()
    return True


def have_reference():
    return os.path.exists(os.path.join(REF_BIN, "Reads_filter"))


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def synth(workdir, args, root="S", threads=4):
    out = subprocess.run([SYNTH] + [str(a) for a in args] + ["--dir", workdir, "--root", root, "--threads",
                         str(threads)], check=True, stdout=subprocess.PIPE, text=True).stdout
    return json.loads(out)


def materialize(name, workdir):
    """Puts the inputs of golden fixture `name` into workdir; returns (root, meta)."""
    src = os.path.join(GOLDEN, name)
    meta = json.load(open(os.path.join(src, "fixture.json")))
    root = meta["root"]
    if "synth_args" in meta:
        synth(workdir, meta["synth_args"] + ["--bps", "0"], root)
    else:
        for f in os.listdir(src):
            if f.startswith(root + ".db") or f.startswith("." + root + "."):
                shutil.copy(os.path.join(src, f), os.path.join(workdir, f))
        with lzma.open(os.path.join(src, root + ".las.xz")) as f, open(os.path.join(workdir, root + ".las"), "wb") as g:
            g.write(f.read())
    assert sha256(os.path.join(workdir, root + ".las")) == meta["las_sha256"], "fixture inputs changed"
    return root, meta


def golden_sums(name):
    sums = {}
    for line in open(os.path.join(GOLDEN, name, "SHA256SUMS")):
        digest, fname = line.split()
        sums[fname[len("out."):]] = digest
    return sums


def stage_cmd(kind, stage, root, prefix, out=None, ini=INI):
    """kind: 'product' | 'oracle' | 'reference'."""
    if kind == "product":
        cmd = [HINGE, stage]
    elif kind == "oracle":
        cmd = [ORACLE, stage]
    else:
        cmd = [os.path.join(REF_BIN, {"filter": "Reads_filter", "maximal": "get_maximal_reads",
                                      "layout": "hinging"}[stage])]
    cmd += ["--db", root, "--las", root + ".las", "-x", prefix, "--config", ini]
    if stage == "layout":
        cmd += ["-o", out or prefix]
    return cmd


def run_stage(kind, stage, workdir, root, prefix, out=None, ini=INI, check=True):
    r = subprocess.run(stage_cmd(kind, stage, root, prefix, out, ini), cwd=workdir, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    if check and r.returncode != 0:
        raise AssertionError("%s %s failed (%d):\n%s" % (kind, stage, r.returncode, r.stdout[-4000:]))
    return r


def first_diff(path_a, path_b, context=2):
    a = open(path_a, errors="replace").read().split("\n")
    b = open(path_b, errors="replace").read().split("\n")
    for i in range(max(len(a), len(b))):
        la = a[i] if i < len(a) else "<EOF>"
        lb = b[i] if i < len(b) else "<EOF>"
        if la != lb:
            return "line %d:\n  got : %s\n  want: %s" % (i + 1, la[:300], lb[:300])
    return "identical"


def assert_same_files(workdir, got_prefix, want_prefix, exts):
    bad = []
    for ext in exts:
        g, w = os.path.join(workdir, got_prefix + "." + ext), os.path.join(workdir, want_prefix + "." + ext)
        assert os.path.exists(w), w
        if not os.path.exists(g):
            bad.append("%s: missing" % ext)
        elif sha256(g) != sha256(w):
            bad.append("%s: %s" % (ext, first_diff(g, w)))
    assert not bad, "output files differ:\n" + "\n".join(bad)


def assert_matches_golden(name, workdir, prefix, exts):
    sums = golden_sums(name)
    bad = []
    for ext in exts:
        g = os.path.join(workdir, prefix + "." + ext)
        if not os.path.exists(g):
            bad.append("%s: missing" % ext)
            continue
        if sha256(g) != sums[ext]:
            want = os.path.join(GOLDEN, name, "expected", "out." + ext)
            bad.append("%s: %s" % (ext, first_diff(g, want) if os.path.exists(want) else "sha256 mismatch"))
    assert not bad, "differs from the reference's golden output (%s):\n%s" % (name, "\n".join(bad))


def split_las(path, base, nparts):
    """LAsplit stand-in: cuts a .las into `nparts` files base.1.las ... at A-read boundaries, about the
    same number of records each (thirdparty/DALIGNER/LAsplit.c:189-194 splits on A-reads too)."""
    import struct

    data = open(path, "rb").read()
    novl, tspace = struct.unpack_from("<qi", data, 0)
    tb = 1 if tspace <= 125 else 2
    pos, starts, areads = 12, [], []
    for _ in range(novl):
        tlen, = struct.unpack_from("<i", data, pos)
        aread, = struct.unpack_from("<i", data, pos + 28)
        starts.append(pos)
        areads.append(aread)
        pos += 40 + tlen * tb
    starts.append(pos)
    cuts = [0]
    for p in range(1, nparts):
        k = max(cuts[-1] + 1, novl * p // nparts)
        while k < novl and areads[k] == areads[k - 1]:
            k += 1
        cuts.append(min(k, novl))
    cuts.append(novl)
    names = []
    for p in range(nparts):
        lo, hi = cuts[p], cuts[p + 1]
        assert hi > lo, "too many parts for this .las"
        name = "%s.%d.las" % (base, p + 1)
        with open(name, "wb") as f:
            f.write(struct.pack("<qi", hi - lo, tspace))
            f.write(data[starts[lo]:starts[hi]])
        names.append(name)
    return names


def stage_cmd_mlas(kind, stage, root, base, prefix, out=None, ini=INI):
    cmd = stage_cmd(kind, stage, root, prefix, out, ini)
    i = cmd.index("--las")
    cmd[i + 1] = base
    return cmd + ["--mlas"]


def run_stage_mlas(kind, stage, workdir, root, base, prefix, out=None, ini=INI, check=True):
    r = subprocess.run(stage_cmd_mlas(kind, stage, root, base, prefix, out, ini), cwd=workdir, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    if check and r.returncode != 0:
        raise AssertionError("%s %s --mlas failed (%d):\n%s" % (kind, stage, r.returncode, r.stdout[-4000:]))
    return r
