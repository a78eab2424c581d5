"""In-tree build of the native libraries (sm_100a only).

  libjsplayer_cuda.so   the product: CUDA kernels + C ABI (include/jsplayer_cuda.h)
  libjsplayer_synth.so  synthetic bitstream encoders (test/bench input generator, plain C)

nvcc cross-compiles without a GPU, so this runs on the CPU-only build container too.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SYNTH = os.path.join(HERE, "synth")
CUDA_LIB = os.path.join(HERE, "libjsplayer_cuda.so")
SYNTH_LIB = os.path.join(HERE, "libjsplayer_synth.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-pthread", "-shared",
]


def _sources(d, exts):
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(exts))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_cuda(force=False, verbose=False):
    srcs = _sources(CSRC, (".cu", ".cpp"))
    deps = srcs + _sources(CSRC, (".cuh", ".h")) + [os.path.join(HERE, "..", "include", "jsplayer_cuda.h")]
    if not force and not _stale(CUDA_LIB, deps):
        return CUDA_LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", CUDA_LIB] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libjsplayer_cuda.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return CUDA_LIB


def build_synth(force=False):
    srcs = _sources(SYNTH, (".c",))
    if not force and not _stale(SYNTH_LIB, srcs):
        return SYNTH_LIB
    cc = os.environ.get("CC", "gcc")
    cmd = [cc, "-O2", "-std=gnu11", "-fPIC", "-shared", "-Wall", "-o", SYNTH_LIB] + srcs + ["-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("gcc failed building libjsplayer_synth.so")
    return SYNTH_LIB


def build_all(force=False, verbose=False):
    return build_cuda(force, verbose), build_synth(force)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
