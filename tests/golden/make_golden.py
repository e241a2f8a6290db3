#!/usr/bin/env python3
"""Regenerates tests/golden/ by running the UNMODIFIED reference binaries
(oracle/_ref/bin, built by oracle/build_ref.sh from /root/reference) on

  dal_small    a real daligner-made fixture: DAZZ_DB `simulator 0.12 -c25. -r3`
               -> fasta2DB -> DBsplit -x500 -s400 -> daligner -> LAsort/LAmerge
               -> DASqv -c25   (278 reads after DBsplit trimming / 11 978 overlaps)
  synth_*      hinge_synth fixtures (tools/hg_synth.cpp), regenerated from the
               parameters in fixture.json at test time; only the reference's
               OUTPUTS are committed, plus the sha256 of the generated inputs.

For every fixture `expected/` holds the reference's output files (small ones
verbatim) and SHA256SUMS lists the digest of every output, large ones included.
Only this container has /root/reference; the GPU box uses the committed files.  dal_small cannot
be re-run against the reference binaries on the GPU box at all: they load the bases (.D.bps),
which are not committed -- re-derive it here (this script) instead.
"""
import hashlib
import json
import lzma
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "bin")
SYNTH = os.path.join(ROOT, "tools", "_build", "hinge_synth")
INI = os.path.join(HERE, "nominal.ini")

OUTPUTS = ["mas", "cmas", "coverage.txt", "repeat.txt", "hinges.txt", "cov.flag", "self.flag", "max",
           "contained.txt", "edges.hinges", "edges.hinges2", "hinge.list", "hgraph", "killed.hinges",
           "edges.skipped", "edges.greedy", "edges.1", "edges.2", "deadends.txt", "garbage.txt"]
KEEP_VERBATIM_BELOW = 200_000

SYNTH_FIXTURES = {
    # name: hinge_synth arguments
    "synth_small": ["--genome", "400000", "--cov", "30", "--seed", "11", "--families", "3"],
    "synth_long": ["--genome", "600000", "--cov", "25", "--seed", "5", "--read-mean", "9000",
                   "--read-sd", "4000", "--read-min", "2000", "--families", "4", "--rep-max", "20000"],
    "synth_noqv": ["--genome", "300000", "--cov", "40", "--seed", "23", "--qv", "0", "--families", "2",
                   "--jitter", "0"],
    # BASELINE configs[4] shape in small: long reads, most pairs reported as two or three local alignments
    "synth_frag": ["--genome", "1500000", "--cov", "40", "--seed", "4321", "--read-mean", "24000",
                   "--read-sd", "8000", "--read-min", "2000", "--frag", "1.2", "--families", "6"],
}


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def run(cmd, cwd):
    subprocess.run(cmd, cwd=cwd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def run_reference(work, root):
    run([os.path.join(REF, "Reads_filter"), "--db", root, "--las", root + ".las", "-x", root, "--config", INI], work)
    run([os.path.join(REF, "get_maximal_reads"), "--db", root, "--las", root + ".las", "-x", root, "--config", INI], work)
    run([os.path.join(REF, "hinging"), "--db", root, "--las", root + ".las", "-x", root, "--config", INI, "-o", root], work)


def collect(work, root, dest):
    exp = os.path.join(dest, "expected")
    shutil.rmtree(exp, ignore_errors=True)
    os.makedirs(exp)
    sums = {}
    for ext in OUTPUTS:
        src = os.path.join(work, root + "." + ext)
        sums[ext] = sha256(src)
        if os.path.getsize(src) < KEEP_VERBATIM_BELOW:
            shutil.copy(src, os.path.join(exp, "out." + ext))
    with open(os.path.join(dest, "SHA256SUMS"), "w") as f:
        for ext in OUTPUTS:
            f.write("%s  out.%s\n" % (sums[ext], ext))


def make_dal_small():
    dest = os.path.join(HERE, "dal_small")
    os.makedirs(dest, exist_ok=True)
    env = dict(os.environ, PATH=REF + ":" + os.environ["PATH"])
    with tempfile.TemporaryDirectory() as work:
        sh = ("simulator 0.12 -c25. -r3 > D.fasta 2>/dev/null && fasta2DB D D.fasta && DBsplit -x500 -s400 D && "
              "daligner D D >/dev/null 2>&1 && LAsort D.D.*.las && LAmerge D D.D.*.S.las && "
              "for f in D.D.*.las; do unlink $f; done && DASqv -c25 D D.las >/dev/null 2>&1")
        subprocess.run(["bash", "-c", sh], cwd=work, env=env, check=True)
        run_reference(work, "D")
        for f in ["D.db", ".D.idx", ".D.qual.anno", ".D.qual.data"]:
            shutil.copy(os.path.join(work, f), os.path.join(dest, f.lstrip(".") if False else f))
        with open(os.path.join(work, "D.las"), "rb") as f, lzma.open(os.path.join(dest, "D.las.xz"), "wb", preset=9) as g:
            g.write(f.read())
        collect(work, "D", dest)
        meta = {"root": "D", "las_sha256": sha256(os.path.join(work, "D.las"))}
        json.dump(meta, open(os.path.join(dest, "fixture.json"), "w"), indent=1)


def make_synth(name, args):
    dest = os.path.join(HERE, name)
    os.makedirs(dest, exist_ok=True)
    with tempfile.TemporaryDirectory() as work:
        out = subprocess.run([SYNTH] + args + ["--dir", work, "--root", "S", "--threads", "3"], check=True,
                             stdout=subprocess.PIPE, text=True).stdout
        info = json.loads(out)
        run_reference(work, "S")
        collect(work, "S", dest)
        meta = {"root": "S", "synth_args": args, "las_sha256": sha256(os.path.join(work, "S.las")),
                "idx_sha256": sha256(os.path.join(work, ".S.idx")), **info}
        json.dump(meta, open(os.path.join(dest, "fixture.json"), "w"), indent=1)


if __name__ == "__main__":
    which = sys.argv[1:] or ["dal_small"] + list(SYNTH_FIXTURES)
    for name in which:
        if name == "dal_small":
            make_dal_small()
        else:
            make_synth(name, SYNTH_FIXTURES[name])
        print("made", name)
