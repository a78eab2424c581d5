"""In-tree build of libjsplayer_synth.so: the synthetic bitstream encoders (plain C, gcc).

Test/bench INPUT GENERATOR -- not part of the product (jsplayer_b200/) and not part of the checker (oracle/)."""
import fcntl
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SYNTH_LIB = os.path.join(HERE, "libjsplayer_synth.so")


def _sources(exts):
    return sorted(os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(exts))


def _digest(deps):
    h = hashlib.sha256()
    for d in sorted(deps):
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _stale(deps):
    try:
        return not os.path.exists(SYNTH_LIB) or open(SYNTH_LIB + ".srchash").read().strip() != _digest(deps)
    except OSError:
        return True


def build(force=False):
    srcs = _sources((".c",))
    deps = srcs + _sources((".h",))
    if not force and not _stale(deps):
        return SYNTH_LIB
    with open(SYNTH_LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or _stale(deps):
                cc = os.environ.get("CC", "gcc")
                tmp = SYNTH_LIB + ".tmp%d" % os.getpid()
                r = subprocess.run([cc, "-O2", "-std=gnu11", "-fPIC", "-shared", "-Wall", "-o", tmp] + srcs + ["-lm"],
                                   capture_output=True, text=True)
                if r.returncode != 0:
                    sys.stderr.write(r.stdout + r.stderr)
                    raise RuntimeError("gcc failed building libjsplayer_synth.so")
                os.replace(tmp, SYNTH_LIB)
                with open(SYNTH_LIB + ".srchash", "w") as fh:
                    fh.write(_digest(deps))
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return SYNTH_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
