"""Build libdifffacto_b200.so (all CUDA sources, sm_100a only) in-tree with nvcc.

    python -m difffacto_b200.build [--force] [-v] [--diag]

`--diag` additionally builds libdifffacto_b200_diag.so: the same sources with -DDFB200_DIAGNOSTICS, which adds the tcgen05
self-tests, the MMA issue-rate microbenchmarks and the phase-timeline hook (include/difffacto_b200_diag.h).  Tests and
tools load it through `_lib.load_diag()`; the product library exports none of it.

The .so is git-ignored but travels to the GPU box with the repo snapshot.  Static cudart: the
library shares the primary context (and the caller's streams) with whatever runtime the host
framework uses.
"""
import argparse
import glob
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libdifffacto_b200.so")
LIB_DIAG = os.path.join(LIBDIR, "libdifffacto_b200_diag.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _digest(flags):
    h = hashlib.sha256()
    inc = os.path.join(PKG, "..", "include")
    for f in sorted(glob.glob(os.path.join(CSRC, "*")) + glob.glob(os.path.join(inc, "*.h"))):
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode() + b"\0" + fh.read())
    h.update(" ".join(flags).encode())
    return h.hexdigest()


def build(force=False, verbose=False, diag=False):
    """Build the product library, or (diag=True) the diagnostic variant; returns the path of the .so."""
    lib = LIB_DIAG if diag else LIB
    objdir = os.path.join(LIBDIR, "diag") if diag else LIBDIR
    flags = NVCC_FLAGS + (["-DDFB200_DIAGNOSTICS"] if diag else [])
    os.makedirs(objdir, exist_ok=True)
    stamp = os.path.join(objdir, "build.sha256")
    dig = _digest(flags)
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return lib
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [_nvcc()] + flags + ["-Xptxas", "-v" if verbose else "-O3", "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {os.path.basename(src)} (rc={p.returncode})\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building difffacto_b200 (see stderr)")
    cmd = [_nvcc(), "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    with open(stamp, "w") as f:
        f.write(dig)
    return lib


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("-v", action="store_true")
    ap.add_argument("--diag", action="store_true", help="also build libdifffacto_b200_diag.so (-DDFB200_DIAGNOSTICS)")
    a = ap.parse_args()
    print(build(a.force, a.v))
    if a.diag:
        print(build(a.force, a.v, diag=True))
