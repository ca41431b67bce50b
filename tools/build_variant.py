"""A/B builds: python tools/build_variant.py NAME -DFLAG[=V] ...  ->  variants/libclipcap_NAME.so

The same sources as tools/build.py with extra nvcc flags (kernel experiments are compiled in behind macros while they
are being measured); select one at run time with CCB_LIB=variants/libclipcap_NAME.so.  variants/ travels to the GPU
box (built .so files are git-ignored, not gpurun-ignored)."""
import concurrent.futures
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import build as B


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    nvcc = B.find_nvcc()
    obj_dir = os.path.join(B.ROOT, "build", "var_" + name)
    os.makedirs(obj_dir, exist_ok=True)
    out_dir = os.path.join(B.ROOT, "variants")
    os.makedirs(out_dir, exist_ok=True)

    def one(src):
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        r = subprocess.run([nvcc] + B.NVCC_FLAGS + flags + ["-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(one, B.sources()))
    out = os.path.join(out_dir, "libclipcap_%s.so" % name)
    r = subprocess.run([nvcc, "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr)
    print(out)


if __name__ == "__main__":
    main()
