"""Per-section cycle counts (clock64) of ScreenPressor P-frame decoding (build with JSP_PROFILE_SECTIONS, see
tools/sp_section_profile.py).  Workload: bench.py's C4 content (1920x1080, 1 I + 31 P, 2 % of the blocks change per
frame), 16 streams."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jsplayer_b200 import BatchDecoder, StreamSpec, CodecType, _lib
import synth
lib = _lib.load()
if not hasattr(lib, "jsp_debug_spp_profile"):
    raise SystemExit("build the library with section timers first:\n"
                     "  JSP_NVCC_EXTRA=-DJSP_PROFILE_SECTIONS python jsplayer_b200/build.py --force")
for ver in (2, 4):
    specs = []
    for i in range(16):
        fr, k, _ = synth.sp_stream(1920, 1080, 32, seed=0xC0DEC4 + i, version=ver, gop=0, change_permille=20)
        specs.append(StreamSpec(CodecType.codec_screenpressor, 1920, 1080, 24, frames=fr, keys=k))
    bd = BatchDecoder(); bd.configure(specs); bd.upload(); bd.run(); bd.sync()
    out = (C.c_ulonglong * 10)()
    lib.jsp_debug_spp_profile(out, 1)
    bd.run(); bd.sync()
    lib.jsp_debug_spp_profile(out, 1)
    v = [int(x) for x in out]
    frames = max(v[7], 1)
    print("version", ver, "P frames", frames, "runs/frame %.0f" % (v[5] / frames), "row pieces/frame %.0f" % (v[6] / frames))
    for n, c in zip(["header + block types", "symbol decodes (P, rgb, N)", "run writes", "sub-rect / motion", "whole frame"], v[:5]):
        print("  %-28s %6.1f%%   %9.0f cycles per frame" % (n, 100.0 * c / max(v[4], 1), c / frames))
    print("  slowest frame %d cycles, mean %.0f; symbols per frame (I included) %.0f" % (v[8], v[4] / frames, bd.symbols() / frames))
    bd.close()
