"""oracle/_ref: the UNMODIFIED reference, byte-compiled -- TEST / BASELINE INFRASTRUCTURE ONLY.

The reference is pure Python: its "build" is CPython bytecode.  `build_ref()` compiles every module of the reference tree FROM
THE SOURCES WHERE THEY LIE (/root/reference, read-only) into sourceless `.pyc` files under oracle/_ref/ (git-ignored, not
gpurun-ignored: like a compiled `.so` it travels to the GPU box, where /root/reference does not exist).  No reference source
text enters the repository.  oracle/ref_harness.py imports the modules from there when the source tree is absent; the only
consumer is `bench.py --impl reference` / its `cpu_baseline` leg, which then time the reference's OWN generation loop
(inference.py:70-148) on the host cores (`cpu_baseline.kind == "reference"`) instead of the oracle's restatement (`"port"`).
Called by `__graft_entry__.build()` in the build container; a no-op wherever the reference tree is absent.
"""
import os
import py_compile
import shutil
import warnings
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SOURCE_ROOT = "/root/reference"


def build_ref(source_root: str = SOURCE_ROOT, out: str = OUT) -> int:
    """Returns the number of modules compiled (0: no reference tree here, whatever is already under oracle/_ref stays)."""
    if not os.path.isdir(os.path.join(source_root, "layers")):
        return 0
    n = 0
    if os.path.isdir(out):
        shutil.rmtree(out)
    for dirpath, dirnames, filenames in os.walk(source_root):
        dirnames[:] = [d for d in dirnames if not d.startswith(".") and d != "__pycache__"]
        for f in filenames:
            if not f.endswith(".py"):
                continue
            src = os.path.join(dirpath, f)
            rel = os.path.relpath(src, source_root)
            dst = os.path.join(out, rel + "c")            # legacy sourceless layout: pkg/mod.pyc beside where mod.py would be
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore", SyntaxWarning)     # the reference's regex literals; not ours to fix
                    py_compile.compile(src, cfile=dst, dfile=rel, doraise=True, optimize=0)
                n += 1
            except py_compile.PyCompileError as e:      # a script of the reference that this interpreter cannot parse is not on the path
                sys.stderr.write("oracle/build_ref: skipped %s (%s)\n" % (rel, e.msg.strip().splitlines()[-1]))
    with open(os.path.join(out, "BUILD_INFO"), "w") as fh:
        fh.write("byte-compiled from %s by oracle/build_ref.py with CPython %s: %d modules\n" % (source_root, sys.version.split()[0], n))
    return n


if __name__ == "__main__":
    print(build_ref())
