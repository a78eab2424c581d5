"""In-tree build of the product library (sm_100a only).

  libjsplayer_cuda.so   CUDA kernels + C ABI (include/jsplayer_cuda.h)

(The synthetic bitstream encoders live in the top-level synth/ package and build themselves, as does the CPU checker.)

nvcc cross-compiles without a GPU, so this runs on the CPU-only build container too.
"""
import fcntl
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
CUDA_LIB = os.path.join(HERE, "libjsplayer_cuda.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-pthread", "-shared",
]


def _sources(d, exts):
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(exts))


def _digest(deps, flags):
    h = hashlib.sha256(" ".join(flags).encode())
    for d in sorted(deps):
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _stale(target, deps, flags=()):
    """Content-based (not mtime-based): a snapshot copied to another machine keeps its built libraries valid."""
    if not os.path.exists(target):
        return True
    try:
        return open(target + ".srchash").read().strip() != _digest(deps, flags)
    except OSError:
        return True


def _stamp(target, deps, flags=()):
    with open(target + ".srchash", "w") as fh:
        fh.write(_digest(deps, flags))


class _Lock:
    """One builder at a time (torchrun starts several ranks that may all find a stale library)."""

    def __init__(self, target):
        self.path = target + ".lock"

    def __enter__(self):
        self.fh = open(self.path, "w")
        fcntl.flock(self.fh, fcntl.LOCK_EX)

    def __exit__(self, *a):
        fcntl.flock(self.fh, fcntl.LOCK_UN)
        self.fh.close()


def build_cuda(force=False, verbose=False):
    srcs = _sources(CSRC, (".cu", ".cpp"))
    deps = srcs + _sources(CSRC, (".cuh", ".h")) + [os.path.join(HERE, "..", "include", "jsplayer_cuda.h")]
    extra = os.environ.get("JSP_NVCC_EXTRA", "").split()          # e.g. -DJSP_PROFILE_SECTIONS (diagnostic builds)
    flags = NVCC_FLAGS + extra
    if not force and not _stale(CUDA_LIB, deps, flags):
        return CUDA_LIB
    with _Lock(CUDA_LIB):
        if not force and not _stale(CUDA_LIB, deps, flags):          # another process built it while we waited
            return CUDA_LIB
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        tmp = CUDA_LIB + ".tmp%d" % os.getpid()
        # one nvcc per translation unit, in parallel (no relocatable device code: the units share nothing on the device),
        # then one link -- the ScreenPressor kernels alone take two minutes
        import tempfile
        from concurrent.futures import ThreadPoolExecutor
        objdir = tempfile.mkdtemp(prefix="jsp_build_")
        cflags = [f for f in flags if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])

        # objects are cached by content (the unit, every header, the flags): touching one kernel recompiles one unit
        cache = os.environ.get("JSP_OBJ_CACHE", os.path.join(tempfile.gettempdir(), "jsp_obj_cache"))
        hdrs = [d for d in deps if d not in srcs]

        def compile_one(src):
            obj = os.path.join(objdir, os.path.basename(src) + ".o")
            key = os.path.join(cache, os.path.basename(src) + "." + _digest([src] + hdrs, cflags)[:24] + ".o")
            if not verbose and os.path.exists(key):
                shutil.copyfile(key, obj)
                return obj, subprocess.CompletedProcess([], 0, "", "")
            r = subprocess.run([nvcc] + cflags + ["-c", "-o", obj, src], capture_output=True, text=True)
            if r.returncode == 0:
                try:
                    os.makedirs(cache, exist_ok=True)
                    shutil.copyfile(obj, key + ".tmp%d" % os.getpid())
                    os.replace(key + ".tmp%d" % os.getpid(), key)
                except OSError:
                    pass
            return obj, r
        with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
            results = list(ex.map(compile_one, srcs))
        log = "".join(r.stdout + r.stderr for _, r in results)
        if any(r.returncode != 0 for _, r in results):
            sys.stderr.write(log)
            raise RuntimeError("nvcc failed building libjsplayer_cuda.so")
        r = subprocess.run([nvcc] + flags + ["-o", tmp] + [o for o, _ in results], capture_output=True, text=True)
        for o, _ in results:
            os.unlink(o)
        os.rmdir(objdir)
        if r.returncode != 0:
            sys.stderr.write(log + r.stdout + r.stderr)
            raise RuntimeError("nvcc failed linking libjsplayer_cuda.so")
        os.replace(tmp, CUDA_LIB)
        _stamp(CUDA_LIB, deps, flags)
        if verbose:
            sys.stderr.write(log)
    return CUDA_LIB


def build_all(force=False, verbose=False):
    return (build_cuda(force, verbose),)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
