"""Build the UNMODIFIED reference CUDA extensions for sm_100a into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  The resulting .so files are the reference's own kernels
(pointnet2_ops._ext, chamfer, emd) compiled from the sources where they lie under
/root/reference; they are used by tests/ (GPU parity: indices bit-exact) and never by
the product path.  Nothing is copied into the repo: oracle/_ref/ is git-ignored but
travels to the GPU box with the snapshot.

Reference sources compiled (read-only):
  pointnet2_ops_lib/pointnet2_ops/_ext-src/src/*.{cpp,cu}   (bindings.cpp:6-19)
  python/difffacto/metrics/chamfer_dist/{chamfer_cuda.cpp,chamfer.cu}
  python/difffacto/metrics/emd/{emd.cpp,emd_cuda.cu}
The reference's own setup.py / JIT arch list ("3.7+PTX;5.0;...") is not used: nvcc 12.9
rejects it, so the arch is forced to 10.0a here.

Usage:  python oracle/build_ref.py [--ref /root/reference]
"""
import argparse
import glob
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def build(ref="/root/reference", verbose=False):
    if not os.path.isdir(ref):
        print(f"[oracle/build_ref] {ref} not present; keeping prebuilt oracle/_ref as is")
        return False
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 4))
    from torch.utils.cpp_extension import load

    S = os.path.join(ref, "pointnet2_ops_lib/pointnet2_ops/_ext-src")
    C = os.path.join(ref, "python/difffacto/metrics/chamfer_dist")
    E = os.path.join(ref, "python/difffacto/metrics/emd")
    jobs = [
        ("ref_pointnet2_ext", glob.glob(S + "/src/*.cpp") + glob.glob(S + "/src/*.cu"), [S + "/include"]),
        ("ref_chamfer", [C + "/chamfer_cuda.cpp", C + "/chamfer.cu"], []),
        ("ref_emd", [E + "/emd.cpp", E + "/emd_cuda.cu"], []),
    ]
    for name, srcs, incs in jobs:
        final = os.path.join(OUT, name + ".so")
        if os.path.exists(final):
            print(f"[oracle/build_ref] {final} exists")
            continue
        bdir = os.path.join("/tmp", "oracle_ref_build", name)
        os.makedirs(bdir, exist_ok=True)
        os.makedirs(OUT, exist_ok=True)
        load(name, sources=srcs, extra_include_paths=incs, extra_cflags=["-O3"],
             extra_cuda_cflags=["-O3"], with_cuda=True, build_directory=bdir,
             verbose=verbose, is_python_module=False)
        shutil.copy(os.path.join(bdir, name + ".so"), final)
        print(f"[oracle/build_ref] built {final}")
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("-v", action="store_true")
    a = ap.parse_args()
    sys.exit(0 if build(a.ref, a.v) else 0)
