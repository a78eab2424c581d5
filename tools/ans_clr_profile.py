"""Where the rANS colour decoder's cycles go (clock64 sections in decodeClr; build with JSP_NVCC_EXTRA=-DJSP_PROFILE_SECTIONS).
Workload: n streams x 1 I frame, 1280x720, v4."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jsplayer_b200 import BatchDecoder, StreamSpec, CodecType, _lib
import synth
lib = _lib.load()
if not hasattr(lib, "jsp_debug_ans2_profile"):
    raise SystemExit("build with JSP_NVCC_EXTRA=-DJSP_PROFILE_SECTIONS first")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
ver = int(sys.argv[2]) if len(sys.argv) > 2 else 4
fr, k, _ = synth.sp_stream(1280, 720, 1, seed=0xC0DEC3, version=ver, gop=0)
specs = [StreamSpec(CodecType.codec_screenpressor, 1280, 720, 24, frames=fr, keys=k) for _ in range(n)]
bd = BatchDecoder(); bd.configure(specs); bd.upload(); bd.run(); bd.sync()
o3 = (C.c_ulonglong * 16)()
lib.jsp_debug_ans2_profile(o3, 1)
bd.run(); bd.sync()
lib.jsp_debug_ans2_profile(o3, 1)
a = [int(x) for x in o3]
nsym = bd.symbols()
print("streams %d symbols %d" % (n, nsym))
tot = 0
for k, nm in enumerate(["cache hit lookup", "Cx4 hit fast path", "kinds 4-6 generic (lane 0)", "raw kinds (None, Cx1-3)", "Cx7 (global)", "cache miss fill"]):
    if a[8 + k]:
        print("  decodeClr %-28s %9d calls  %6.0f cycles per call  %5.1f%% of colour time" % (nm, a[8 + k] // n, a[k] / a[8 + k], 0))
        tot += a[k]
print("  colour cycles per symbol (all symbols): %.0f" % (tot / nsym))
bd.close()
