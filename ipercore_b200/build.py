"""Build recipe for libiper_b200.so: plain nvcc, sm_100a only, in-tree (the .so travels with the gpurun snapshot).

No torch headers are involved: the library's boundary is the C ABI in include/iper_b200.h.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INC = os.path.abspath(os.path.join(HERE, "..", "include"))
SO = os.path.join(HERE, "libiper_b200.so")
SOURCES = ["api.cu", "raster.cu", "conv_tc.cu", "ops.cu", "lbs.cu", "source.cu", "generator.cu", "train.cu", "train_ops.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-I", INC, "-I", CSRC, "--expt-relaxed-constexpr"]


def _newer(a, b):
    return not os.path.exists(b) or os.path.getmtime(a) > os.path.getmtime(b)


def build(verbose=False, force=False):
    objs = []
    deps = [os.path.join(CSRC, "common.cuh"), os.path.join(INC, "iper_b200.h")]
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(s, o) or any(_newer(d, o) for d in deps):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("==== %s ====\n%s\n" % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or not os.path.exists(SO):
        subprocess.check_call([NVCC, "-shared", "-o", SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                  "-lcudart"])
    return SO


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
