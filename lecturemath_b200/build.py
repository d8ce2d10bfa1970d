"""Builds lecturemath_b200/libaccessmath_b200.so in-tree with nvcc for sm_100a (no JIT cache: the built
.so travels to the GPU box with the repo snapshot).   python -m lecturemath_b200.build [--force]"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libaccessmath_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "accessmath_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, variant=None, defines=()):
    """variant/defines: tuning experiments only (e.g. variant="e16", defines=["-DEPI_WARPS=16"]) -> a second library
    libaccessmath_b200_<variant>.so that AM_B200_LIB can point the binding at."""
    out = OUT if not variant else OUT[:-3] + "_" + variant + ".so"
    if not force and not variant and not _stale():
        return out
    objs = []
    bdir = os.path.join(HERE, "build" + ("_" + variant if variant else ""))
    os.makedirs(bdir, exist_ok=True)
    for src in sources():
        obj = os.path.join(bdir, os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC] + FLAGS + list(defines) + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        subprocess.check_call(cmd)
        objs.append(obj)
    subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs + ["-lcudart_static", "-ldl", "-lpthread", "-lrt"])
    return out


if __name__ == "__main__":
    args = [a for a in sys.argv[1:]]
    variant = None
    if "--variant" in args:
        variant = args[args.index("--variant") + 1]
    print(build(force="--force" in args, verbose="-v" in args, variant=variant, defines=[a for a in args if a.startswith("-D")]))
