"""Cycles per symbol of the ScreenPressor entropy kernels against the number of warps resident per SM (same data in every
warp: one 1280x720 I frame replicated), per coder.  A latency-bound kernel with independent warps should hold its cycles
per symbol as warps are added; if they rise, co-resident warps interfere (instruction cache, shared-memory pipe, L1).

    python tools/sp_coresidency.py [--frames-per-sm 1,2,3,4,5,6,7] [--sms 148] > profiles/rNN_sp_coresidency.txt

Prints one line per (coder, warps per SM): launch ms, symbols per stream, cycles per symbol per stream (ms x SM clock /
symbols of ONE stream -- all streams decode the same frame, so with one wave the launch lasts as long as one stream).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jsplayer_b200 import BatchDecoder, StreamSpec, CodecType, _lib       # noqa: E402
import synth                                                               # noqa: E402


def sm_clock_mhz():
    try:
        import subprocess
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.max.sm", "--format=csv,noheader,nounits", "-i", "0"],
                             capture_output=True, text=True, timeout=20).stdout.strip().splitlines()[0]
        return float(out)
    except Exception:
        return 1965.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames-per-sm", default="1,2,3,4,5,6,7")
    ap.add_argument("--sms", type=int, default=148)
    ap.add_argument("--size", default="1280x720")
    ap.add_argument("--versions", default="2,4")
    ap.add_argument("--pframes", type=int, default=0, help="append this many P frames per stream (C4-like content)")
    a = ap.parse_args()
    w, h = (int(v) for v in a.size.split("x"))
    mhz = sm_clock_mhz()
    _lib.require_gpu()
    print("# %s, %d SMs, SM clock %.0f MHz (max; cycles assume it), %d P frames per stream" % (a.size, a.sms, mhz, a.pframes))
    print("# coder warps/SM streams launch_ms symbols_per_stream cycles_per_symbol Msymbols_per_s")
    for ver in (int(v) for v in a.versions.split(",")):
        frames, keys, _ = synth.sp_stream(w, h, 1 + a.pframes, seed=0xC0DEC3, version=ver, gop=0, change_permille=40)
        for k in (int(v) for v in a.frames_per_sm.split(",")):
            n = a.sms * k
            specs = [StreamSpec(CodecType.codec_screenpressor, w, h, 24, frames=frames, keys=keys) for _ in range(n)]
            bd = BatchDecoder()
            bd.configure(specs)
            bd.upload()
            ms, kms, cnt = bd.time_runs(warmup=1, iters=3, flush_l2=True)
            ms /= 3.0
            nsym = bd.symbols()
            per = nsym / n
            print("v%d %d %d %.3f %.0f %.1f %.1f" % (ver, k, n, ms, per, ms * 1e-3 * mhz * 1e6 / per, nsym / ms / 1e3), flush=True)
            bd.close()


if __name__ == "__main__":
    main()
