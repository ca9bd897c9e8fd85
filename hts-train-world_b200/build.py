"""Build recipe for libworld_b200.so (in-tree, sm_100a only).

  python hts-train-world_b200/build.py [--force]

Every csrc/*.cu is compiled with
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3
into build/*.o (in parallel) and linked into hts-train-world_b200/libworld_b200.so.
nvcc cross-compiles without a GPU, so this runs in the CPU-only container too.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libworld_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-DWB_BUILD",
         "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v"]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    inc = os.path.join(HERE, "..", "include")
    for root, _, files in os.walk(inc):
        hs += [os.path.join(root, f) for f in files]
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, obj, log):
    cmd = [NVCC] + FLAGS + ["-I", os.path.join(HERE, "..", "include"), "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr[-4000:]))
    return obj


def build(force=False, verbose=False, variant=None, defines=()):
    """variant / defines: an experimental build with extra -D flags into libworld_b200_<variant>.so
    (loaded instead of the default library when WB200_LIB points at it: A/B measurements)."""
    global FLAGS
    bdir, lib = BUILD, LIB
    flags_saved = FLAGS
    if variant:
        bdir = os.path.join(BUILD, variant)
        lib = os.path.join(HERE, "libworld_b200_%s.so" % variant)
        FLAGS = FLAGS + ["-D" + d for d in defines]
    try:
        return _build(force, verbose, bdir, lib)
    finally:
        FLAGS = flags_saved


def _build(force, verbose, BUILD, LIB):
    os.makedirs(BUILD, exist_ok=True)
    srcs = _sources()
    dep_m = _deps_mtime()
    jobs = []
    objs = []
    for s in srcs:
        o = os.path.join(BUILD, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), dep_m):
            jobs.append((s, o, o[:-2] + ".log"))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for _ in ex.map(lambda j: _compile(*j), jobs):
                pass
    if jobs or not os.path.exists(LIB):
        # -Bsymbolic-functions: calls between the library's own exported functions (the helpers of
        # wb_compat.cu call interp1Q, fft_execute, ...) bind inside the library, so a host program
        # that happens to define a function of the same name cannot interpose them
        cmd = [NVCC, "-shared", "-Xlinker", "-Bsymbolic-functions", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    if verbose:
        print("built", LIB, "(%d compiled)" % len(jobs))
    return LIB


if __name__ == "__main__":
    # python build.py [--force] [--variant NAME -DFLAG ...]
    var = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    build(force="--force" in sys.argv, verbose=True, variant=var, defines=[a[2:] for a in sys.argv if a.startswith("-D")])
