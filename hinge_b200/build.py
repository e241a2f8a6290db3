"""Builds libhinge_b200.so (CUDA kernels + C ABI + host drivers) and the `hinge`
front-end in-tree with nvcc for sm_100a.  No JIT, no torch dependency: the
artefacts under hinge_b200/_build travel to the GPU box with the snapshot."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_build")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "--extended-lambda"]
CU = ["hg_filter.cu", "hg_filter_flat.cu", "hg_capi.cu", "hg_layout.cu", "hg_capi_layout.cu"]
CPP = ["hg_io.cpp", "hg_host.cpp", "hg_host_layout.cpp"]


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout)
        raise RuntimeError("build step failed: " + cmd[0])
    return r.stdout


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    deps = list(sources) + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(HERE, "..", "include", "hinge_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=False, force=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(OUT, exist_ok=True)
    objs = []
    for f in CU + CPP:
        src = os.path.join(CSRC, f)
        if not os.path.exists(src):
            continue
        obj = os.path.join(OUT, f + ".o")
        if force or _stale(obj, [src]):
            flags = list(NVCC_FLAGS)
            if verbose and f.endswith(".cu"):
                flags += ["-Xptxas", "-v"]
            out = _run([nvcc] + flags + ["-c", src, "-o", obj])
            if verbose:
                print(out)
        objs.append(obj)
    lib = os.path.join(OUT, "libhinge_b200.so")
    if force or _stale(lib, objs):
        _run([nvcc, "-shared", "-o", lib] + objs + ["-cudart", "static"])
    exe = os.path.join(OUT, "hinge")
    cli = os.path.join(CSRC, "hg_cli.cpp")
    if force or _stale(exe, [cli, lib]):
        _run(["g++", "-O2", "-std=c++17", cli, "-o", exe, "-L" + OUT, "-lhinge_b200",
              "-Wl,-rpath,$ORIGIN"])
    for name in ("Reads_filter", "get_maximal_reads", "hinging"):
        link = os.path.join(OUT, name)
        if not os.path.islink(link):
            if os.path.exists(link):
                os.remove(link)
            os.symlink("hinge", link)
    return lib


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
