"""Build libfzb200.so in-tree with nvcc for sm_100a (no JIT cache, no torch dependency).

    python -m frankenz_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libfzb200.so")
# (source, extra flags).  The float64 kernels are compiled without FMA contraction so that their
# arithmetic is the IEEE sequence numpy executes (e.g. d - (d/m)*m == 0 for a one-band fit).
SOURCES = [("fzb_api.cu", []), ("fzb_generic.cu", ["-fmad=false"]), ("fzb_fast.cu", []), ("fzb_sweep_tc.cu", []), ("fzb_knn.cu", ["-fmad=false"]), ("fzb_knn_tc.cu", []),
           ("fzb_summarize.cu", ["-fmad=false"]), ("fzb_shard.cu", ["-fmad=false"]), ("fzb_prune.cu", []), ("fzb_samplers.cu", ["-fmad=false"])]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "frankenz_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(force=False, verbose=False):
    """Compile every CUDA source of the package into frankenz_b200/lib/libfzb200.so."""
    if not force and not is_stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for src, extra in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + out)
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
