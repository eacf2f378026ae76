"""In-tree build of libfdms.so (hand-written CUDA for sm_100a behind a C ABI).

    python -m face_detection_multi_scale_b200.build [--force]

nvcc cross-compiles without a GPU, so this runs in the build container; the resulting .so is
git-ignored but travels to the GPU box with the repo snapshot.  There is no fallback: if the
library is missing and cannot be built, importing the ops raises.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfdms.so")
OBJ_DIR = os.path.join(HERE, "csrc", "_obj")
SOURCES = ["ctx.cu", "pyramid.cu", "pyramid_strip.cu", "decode.cu", "nms.cu", "merge.cu", "eval.cu", "tta.cu", "jpeg.cu", "jpeg_host.cu"]
HEADERS = ["common.cuh", "pyramid.cuh", "jpeg_host.cuh", "nms.cuh", "merge_rows.cuh", os.path.join("..", "..", "include", "fdms.h")]

# -fmad=false: every fp32 rounding the reference performs is kept (fused ops are explicit).
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "--fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-diag-suppress", "177"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libfdms.so cannot be built (there is no CPU fallback)")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False):
    """Compile csrc/*.cu into libfdms.so if missing or older than its sources. Returns the path.
    Concurrent callers (the ranks of a torchrun job on a box whose snapshot lost the file times) serialise on a
    lock file; the ones that waited find the library fresh and return."""
    if not force and not _stale():
        return LIB
    import fcntl
    os.makedirs(OBJ_DIR, exist_ok=True)
    with open(os.path.join(OBJ_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():
                return LIB
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force, verbose):
    try:
        nvcc = _nvcc()
    except RuntimeError:
        if os.path.exists(LIB) and not force:  # e.g. a deployment box without the toolkit: use the shipped build
            return LIB
        raise
    os.makedirs(OBJ_DIR, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB + ".tmp.%d" % os.getpid()
    r = subprocess.run([nvcc, "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                              "-cudart", "static"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    p = build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
