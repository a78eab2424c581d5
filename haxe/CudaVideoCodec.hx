package ;
// CudaVideoCodec.hx -- an IVideoCodec (reference src/IVideoCodec.hx:16-29) whose every member forwards to
// libjsplayer_cuda.  On the hxcpp target the reference's typed arrays are replaced by haxe.io.Bytes-backed
// views (the reference imports js.lib.Int32Array / Uint8Array, which exist only on the JS target; the two
// typedefs below are the shim a maintainer adds -- see INTEGRATION.md).
// Shipped as source only (no haxe toolchain in this image).
#if cpp
import cpp.NativeArray;
import cpp.Pointer;

typedef Int32Array = Array<cpp.Int32>;     // contiguous on hxcpp; NativeArray.address gives the raw pointer
typedef Uint8Array = Array<cpp.UInt8>;

class CudaVideoCodec implements IVideoCodec {
    var h : cpp.RawPointer<JsplayerCuda.JspDecNative>;
    var buffers : Map<Int, Int32Array>;    // address (low bits) -> the caller's array, to hand the same object back
    var prev : Int32Array;

    // codec: 0 ScreenPressor, 1 MSVideo1 RGB555, 2 MSVideo1 8-bit (VideoData.hx:75-80)
    public function new(codec:Int, X:Int, Y:Int, bpp:Int, ?palette:Uint8Array, device:Int = -1) {
        var pp : cpp.RawConstPointer<cpp.UInt8> = palette == null ? null : cast NativeArray.address(palette, 0).constRaw;
        h = JsplayerCuda.Jsp.create(codec, X, Y, bpp, pp, palette == null ? 0 : palette.length, device);
        if (h == null) throw "jsp_create: " + JsplayerCuda.Jsp.lastError().toString();
        buffers = new Map();
    }

    public function Preinit(insignificant_lines:Int):Void { JsplayerCuda.Jsp.preinit(h, insignificant_lines); }
    public function PreviousFrame():Int32Array { return prev; }
    public function IsKeyFrame(data:Uint8Array):Bool {
        return JsplayerCuda.Jsp.isKeyFrame(h, cast NativeArray.address(data, 0).constRaw, data.length) != 0;
    }
    public function State():DecoderState { return toState(JsplayerCuda.Jsp.stateOf(h)); }

    public function DecompressI(src:Uint8Array, dst:Int32Array):DecoderState {
        var st = JsplayerCuda.Jsp.decompressI(h, cast NativeArray.address(src, 0).constRaw, src.length, cast NativeArray.address(dst, 0).raw);
        if (st == 0) prev = dst;
        return toState(st);
    }
    public function ContinueI():DecoderState { return toState(JsplayerCuda.Jsp.continueI(h)); }

    public function DecompressP(src:Uint8Array, dst:Int32Array):PFrameResult {
        var r = JsplayerCuda.Jsp.decompressP(h, cast NativeArray.address(src, 0).constRaw, src.length, cast NativeArray.address(dst, 0).raw);
        var dstRaw : cpp.RawPointer<cpp.Int32> = cast NativeArray.address(dst, 0).raw;
        if (r.data_pnt == null) return { data_pnt: null, significant_changes: false };
        if (r.data_pnt == dstRaw) prev = dst;                       // the picture changed: dst is retained
        return { data_pnt: prev, significant_changes: r.significant_changes != 0 };
    }
    public function NeedsIndex():Bool { return JsplayerCuda.Jsp.needsIndex(h) != 0; }
    public function StopAndClean():Void { JsplayerCuda.Jsp.stopAndClean(h); prev = null; }

    static inline function toState(v:Int):DecoderState {
        return switch (v) { case 0: zero_state; case 1: in_progress; default: error_occured; };
    }
}
#end
