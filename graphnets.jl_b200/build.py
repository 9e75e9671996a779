"""In-tree build of libgnb200.so (the C-ABI CUDA library) for sm_100a.

    python graphnets.jl_b200/build.py [--force]

nvcc cross-compiles without a GPU.  The .so lands next to this file, is git-ignored and travels to
the GPU box with the gpurun snapshot."""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libgnb200.so")
BUILD = os.path.join(HERE, "build")
SOURCES = ["model.cu", "lower.cu", "fp32.cu", "tc.cu", "graphrows.cu", "smallk.cu", "tc_edge.cu", "tc_gemm.cu", "loss.cu", "train.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas=-v"] + os.environ.get("GNB_EXTRA_NVCC_FLAGS", "").split()


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for fn in sorted(os.listdir(root)):
            if fn.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, fn), "rb") as f:
                    h.update(fn.encode())
                    h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


# Test-only variants of the library: {name: (sources recompiled, extra flags)}; every other object is shared with the
# product build.  `oldproj` keeps round 1's k_tc_proj barrier protocol so that tests/test_gpu_watchdog.py can reproduce
# its dead-lock (GNB_LIB_VARIANT=oldproj selects it in _lib.py); the product never loads a variant.
VARIANTS = {"oldproj": (["tc.cu"], ["-DGNB_OLD_PROJ_PROTOCOL"]),
            "timing": (["tc_edge.cu"], ["-DGNB_TC_TIMING"])}      # clock64 phase stamps of the fused kernel (tools/edge_timing.py)
# development only: GNB_AB_FLAGS="-DX=1" python build.py builds libgnb200_ab.so with those flags on the tensor-path sources, for
# same-box A/B runs (tools/ab_edge.sh ab)
# several at once: GNB_AB_FLAGS="ab1:-DX=1;ab2:-DY=1 -DZ" builds libgnb200_ab1.so, libgnb200_ab2.so
if os.environ.get("GNB_AB_FLAGS"):
    for i, spec in enumerate(os.environ["GNB_AB_FLAGS"].split(";")):
        name, _, flags = spec.partition(":") if ":" in spec else ("ab" if i == 0 else "ab%d" % i, "", spec)
        VARIANTS[name.strip()] = (["tc_edge.cu", "tc.cu", "tc_gemm.cu", "fp32.cu", "smallk.cu", "graphrows.cu"], flags.split())


def variant_path(name):
    return os.path.join(HERE, "libgnb200_%s.so" % name)


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(BUILD, "stamp")
    dig = _digest()
    outs = [OUT] + [variant_path(v) for v in VARIANTS]
    if not force and all(os.path.exists(o) for o in outs) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT

    def cc(job):
        src, tag, extra = job
        obj = os.path.join(BUILD, src.replace(".cu", tag + ".o"))
        cmd = [NVCC] + FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose and not tag:
            sys.stderr.write(r.stderr)
        return obj

    jobs = [(s, "", []) for s in SOURCES]
    for v, (srcs, extra) in VARIANTS.items():
        jobs += [(s, "_" + v, extra) for s in srcs]
    with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
        objs = dict(zip([(j[0], j[1]) for j in jobs], ex.map(cc, jobs)))

    def link(out, tag, srcs):
        # cudart is linked statically (nvcc default) and the driver API is resolved at run time via
        # cudaGetDriverEntryPoint, so the .so loads on a box without libcuda (symbol checks on CPU).
        cmd = [NVCC, "-shared", "-o", out] + [objs[(s, tag if s in srcs else "")] for s in SOURCES]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))

    link(OUT, "", [])
    for v, (srcs, _) in VARIANTS.items():
        link(variant_path(v), "_" + v, srcs)
    with open(stamp, "w") as f:
        f.write(dig)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
