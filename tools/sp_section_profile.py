"""Per-section cycle counts (clock64) of the ScreenPressor entropy kernels: where a symbol's ~800-1000 cycles go.

ncu's stall sampling is flat for these kernels (one warp per scheduler, every instruction waits a little), so the I-frame
loop and the decoders carry optional timers (JSP_PROFILE_SECTIONS).  Workload: 32 streams x 4 I frames, 1280x720, both
coders.  Results of round 1: profiles/r01_sp_section_profile.txt, discussion in DESIGN.md 4.3.
"""
import ctypes as C, numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jsplayer_b200 import BatchDecoder, StreamSpec, CodecType, _lib
import synth
lib = _lib.load()
if not hasattr(lib, "jsp_debug_sp_profile"):
    raise SystemExit("build the library with section timers first:\n"
                     "  JSP_NVCC_EXTRA=-DJSP_PROFILE_SECTIONS python jsplayer_b200/build.py --force\n"
                     "(and rebuild without the flag afterwards: the timers cost ~40 cycles per section)")
for ver in (2, 4):
    specs = []
    for i in range(32):
        fr, k, _ = synth.sp_stream(1280, 720, 4, seed=0xC0DEC3 + i, version=ver, gop=1, change_permille=40)
        specs.append(StreamSpec(CodecType.codec_screenpressor, 1280, 720, 24, frames=fr, keys=k))
    bd = BatchDecoder(); bd.configure(specs); bd.upload(); bd.run(); bd.sync()
    out = (C.c_ulonglong * 8)()
    lib.jsp_debug_sp_profile(out, 1)
    bd.run(); bd.sync()
    lib.jsp_debug_sp_profile(out, 1)
    nsym = bd.symbols()
    v = [int(x) for x in out]
    runs = v[5]
    print("version", ver, "warps 128, symbols", nsym, "runs", runs)
    names = ["decodeP", "decode_rgb(3 symbols)", "decodeN", "segment writes", "whole loop"]
    for n, c in zip(names, v[:5]):
        print("  %-24s %6.1f%% of loop   %7.0f cycles per run" % (n, 100.0 * c / v[4], c / runs))
    print("  cycles per symbol (loop) %.0f" % (v[4] / nsym))
    if ver != 2:
        o3 = (C.c_ulonglong * 16)()
        lib.jsp_debug_ans_profile(o3, 1)
        a = [int(x) for x in o3]
        for k, n in enumerate(["cache hit lookup", "Cx4 hit fast path", "kinds 4-6 generic (lane 0)", "raw kinds (None, Cx1-3)", "Cx7 (global)", "cache miss fill"]):
            if a[8 + k]:
                print("    decodeClr %-28s %9d calls  %6.0f cycles per call" % (n, a[8 + k] // 2, a[k] / a[8 + k]))
    if ver == 2:
        o2 = (C.c_ulonglong * 8)()
        lib.jsp_debug_rc_profile(o2, 1)
        w = [int(x) for x in o2]
        nbig = nsym - runs          # decode_big calls = symbols - decodeP calls
        for n, c in zip(["load table/row", "division", "ballot + search", "shuffles", "consume/renorm", "update + store"], w[:6]):
            print("    decode_big %-18s %6.0f cycles per call" % (n, c / nbig / 2))   # two runs accumulated
    bd.close()
