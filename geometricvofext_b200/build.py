"""In-tree build of the CUDA library (sm_100a only; nvcc cross-compiles without a GPU).

    python -m geometricvofext_b200.build [--force] [--verbose]

Seven translation units are compiled in parallel (the host/streaming TU + one per
polyhedron-capacity variant of the geometry kernels) and linked into
geometricvofext_b200/lib/libsvof_b200.so.

-fmad=false is a CORRECTNESS flag here, not a tuning knob: the parity contract
(bit-exact interface-cell set / cut-face topology, alpha within 1e-12 per step)
needs the reference's operation order without DFMA contraction.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HEADERS = [os.path.join(CSRC, f) for f in ("svof_kernels.cuh", "svof_geom_kernels.cuh", "svof_geom.cuh", "svof_math.cuh", "svof_plic_group.cuh", "svof_plic_warp.cuh")] + \
          [os.path.join(HERE, "..", "include", "svof.h")]
_TAG = os.environ.get("SVOF_BUILD_TAG", "")   # variants build beside the product: lib/libsvof_b200<tag>.so
OUT = os.path.join(HERE, "lib", "libsvof_b200%s.so" % _TAG)
OBJ = os.path.join(HERE, "lib", "obj" + _TAG)

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-Xcompiler", "-fPIC"] + \
             (["-DSV_BOUND_STATS"] if os.environ.get("SVOF_BOUND_STATS") else []) + \
             (["-DSV_DENSE_UNROLL=" + os.environ["SVOF_DENSE_UNROLL"]] if os.environ.get("SVOF_DENSE_UNROLL") else []) + \
             os.environ.get("SVOF_EXTRA_DEFS", "").split()   # kernel-variant experiments (with SVOF_BUILD_TAG)

UNITS = [("svof_b200", "svof_b200.cu", []), ("svof_decomp", "svof_decomp.cpp", [])] + \
        [("svof_inst%d" % v, "svof_inst.cu", ["-DSV_VARIANT=%d" % v]) for v in range(5)]


def _stale(target, deps):
    return (not os.path.exists(target)) or any(os.path.getmtime(d) > os.path.getmtime(target) for d in deps)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    jobs = []
    for name, src, defs in UNITS:
        obj = os.path.join(OBJ, name + ".o")
        srcp = os.path.join(CSRC, src)
        if force or _stale(obj, [srcp] + HEADERS):
            cmd = [nvcc] + NVCC_FLAGS + defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, srcp]
            jobs.append((name, cmd))
    if jobs:
        def run(job):
            name, cmd = job
            p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            return name, p.returncode, p.stdout
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            results = list(ex.map(run, jobs))
        log = os.path.join(OBJ, "build.log")
        with open(log, "w") as f:
            for name, rc, out in results:
                f.write("==== %s (rc=%d)\n%s\n" % (name, rc, out))
        for name, rc, out in results:
            if rc != 0:
                sys.stderr.write(out)
                raise RuntimeError("nvcc failed for %s" % name)
    objs = [os.path.join(OBJ, name + ".o") for name, _, _ in UNITS]
    if force or jobs or _stale(OUT, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-cudart", "static", "-Xlinker", "-Bsymbolic", "-Xcompiler", "-pthread", "-o", OUT] + objs + ["-ldl"]
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
