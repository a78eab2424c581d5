"""Where the cycles of the second-generation I-frame kernel go (clock64 sections, build with JSP_NVCC_EXTRA=-DJSP_SP2_PROF).
Workload: n streams x 1 I frame, 1280x720, range coder."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jsplayer_b200 import BatchDecoder, StreamSpec, CodecType, _lib
import synth
lib = _lib.load()
if not hasattr(lib, "jsp_debug_sp2_profile"):
    raise SystemExit("build with JSP_NVCC_EXTRA=-DJSP_SP2_PROF first")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
ver = int(sys.argv[2]) if len(sys.argv) > 2 else 2
fr, k, _ = synth.sp_stream(1280, 720, 1, seed=0xC0DEC3, version=ver, gop=0)
specs = [StreamSpec(CodecType.codec_screenpressor, 1280, 720, 24, frames=fr, keys=k) for _ in range(n)]
bd = BatchDecoder(); bd.configure(specs); bd.upload(); bd.run(); bd.sync()
out = (C.c_ulonglong * 16)()
lib.jsp_debug_sp2_profile(out, 1)
bd.run(); bd.sync()
lib.jsp_debug_sp2_profile(out, 1)
v = [int(x) for x in out]
nsym = bd.symbols()
runs, crun = v[6], v[7]
print("streams %d coder v%d symbols %d runs %d colour runs %d" % (n, ver, nsym, runs, crun))
names = ["decodeP", "decode_rgb", "decodeN", "push", "drain wait", "E whole loop", None, None, "R wait", "R run write", "R whole loop", "loop overhead"]
for i, nm in enumerate(names):
    if nm:
        print("  %-14s %12d cycles  %7.1f per run  %5.1f%% of E loop" % (nm, v[i], v[i] / max(1, runs), 100.0 * v[i] / max(1, v[5])))
print("  decode_rgb per colour run %.1f; E cycles per symbol %.1f" % (v[1] / max(1, crun), v[5] / max(1, nsym)))
bd.close()
