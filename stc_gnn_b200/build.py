"""In-tree build of libstc_b200.so (hand-written sm_100a kernels + the C ABI of include/stc_b200.h).

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box.

Safe with one process per GPU: the digest is taken over package-relative paths (a checkout at another
path does not look stale), the whole build runs under an exclusive file lock, objects and the library are
written to temporary names and moved into place with os.replace(), so a concurrent dlopen never sees a
half-written file.  Sources are compiled one object per .cu, in parallel, and only the objects whose
inputs changed are rebuilt.
"""
import fcntl
import glob
import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HEADER = os.path.join(HERE, "..", "include", "stc_b200.h")
LIB = os.path.join(HERE, "libstc_b200.so")
STAMP = os.path.join(HERE, ".libstc_b200.stamp")
OBJDIR = os.path.join(HERE, "build")
LOCK = os.path.join(HERE, ".libstc_b200.lock")

NVCC_COMPILE = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
]
NVCC_LINK = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [HEADER]


def _hash_files(h, paths):
    for p in paths:
        with open(p, "rb") as f:
            h.update(os.path.relpath(p, HERE).replace(os.sep, "/").encode())
            h.update(f.read())


def _digest():
    h = hashlib.sha256(" ".join(NVCC_COMPILE + NVCC_LINK).encode())
    _hash_files(h, _sources() + _headers())
    return h.hexdigest()


def _object_digest(src):
    h = hashlib.sha256(" ".join(NVCC_COMPILE).encode())
    _hash_files(h, [src] + _headers())
    return h.hexdigest()


def _up_to_date(digest):
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == digest


def _compile_one(nvcc, src, verbose):
    obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
    want = _object_digest(src)
    tag = obj + ".sha"
    if os.path.exists(obj) and os.path.exists(tag):
        with open(tag) as f:
            if f.read().strip() == want:
                return obj, ""
    tmp = f"{obj}.{os.getpid()}.tmp"
    cmd = [nvcc] + NVCC_COMPILE + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", tmp, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    os.replace(tmp, obj)
    with open(tag, "w") as f:
        f.write(want)
    return obj, res.stdout + res.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into libstc_b200.so when sources changed. Returns the library path."""
    digest = _digest()
    if not force and _up_to_date(digest):
        return LIB
    nvcc = _nvcc()
    if nvcc is None:
        if os.path.exists(LIB):
            return LIB  # GPU box without a toolkit on PATH: use the shipped build
        raise RuntimeError("nvcc not found and libstc_b200.so has not been built")
    with open(LOCK, "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)       # one builder at a time (ranks of one job share this directory)
        try:
            if not force and _up_to_date(digest):   # another process built it while we waited
                return LIB
            os.makedirs(OBJDIR, exist_ok=True)
            if force:
                for p in glob.glob(os.path.join(OBJDIR, "*.sha")):
                    os.remove(p)
            with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
                results = list(pool.map(lambda s: _compile_one(nvcc, s, verbose), _sources()))
            tmp = f"{LIB}.{os.getpid()}.tmp"
            cmd = [nvcc] + NVCC_LINK + ["-o", tmp] + [o for o, _ in results]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
            os.replace(tmp, LIB)
            if verbose:
                print("".join(log for _, log in results) + res.stdout + res.stderr)
            with open(STAMP + ".tmp", "w") as f:
                f.write(digest)
            os.replace(STAMP + ".tmp", STAMP)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
