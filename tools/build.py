"""Builds clip-image-captioning_b200/libclipcap_b200.so (sm_100a only) with nvcc.

    python tools/build.py [--force] [--verbose]

Objects go to build/ (git-ignored); the shared library stays in-tree next to the Python host code so that it
travels to the GPU box with the repo snapshot.  A source / header newer than the library triggers a rebuild.
"""
import argparse
import concurrent.futures
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "clip-image-captioning_b200")
CSRC = os.path.join(PKG, "csrc")
OUT = os.path.join(PKG, "libclipcap_b200.so")
OBJ_DIR = os.path.join(ROOT, "build", "obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(ROOT, "include", "clipcap_b200.h"))
    return hs


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(p) <= t for p in sources() + headers() + [os.path.abspath(__file__)])


def compile_one(nvcc, src, verbose):
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
    hdr_t = max(os.path.getmtime(h) for h in headers())
    if os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_t, os.path.getmtime(__file__)):
        return obj
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build(force=False, verbose=False, tuning=False):
    """tuning=True compiles the per-CTA timeline stamps in (-DCCB_TUNING: tools/trace_*.py, tools/ab_mega.py need them)."""
    if tuning:
        force = True
        NVCC_FLAGS.append("-DCCB_TUNING")
    if not force and up_to_date():
        return OUT
    nvcc = find_nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(lambda s: compile_one(nvcc, s, verbose), sources()))
    # the driver entry point for tensor maps is resolved at run time (cudaGetDriverEntryPoint): no -lcuda
    cmd = [nvcc, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return OUT


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--tuning", action="store_true", help="compile the timeline stamps in (-DCCB_TUNING)")
    a = ap.parse_args()
    print(build(a.force, a.verbose, a.tuning))
