package ;
// JsplayerCuda.hx -- hxcpp externs for libjsplayer_cuda (include/jsplayer_cuda.h).
//
// Shipped as source: this image has no haxe / hxcpp toolchain (`haxe`, `haxelib`, `node` are absent), so the
// file has not been compiled here.  It binds exactly the C-ABI entry points the header declares; every extern
// below names the IVideoCodec member (reference src/IVideoCodec.hx:16-29) it stands in for.
// Build: add haxe/Build.xml's <target> snippet to the hxcpp build (links -ljsplayer_cuda) and compile the
// reference with `-cpp out -D jsplayer_cuda` (INTEGRATION.md).
import cpp.ConstPointer;
import cpp.Pointer;
import cpp.RawConstPointer;
import cpp.RawPointer;
import cpp.UInt8;
import cpp.Int32;

@:include("jsplayer_cuda.h")
@:native("jsp_dec")
extern class JspDecNative {}

@:include("jsplayer_cuda.h")
@:structAccess
@:native("jsp_pframe_result")
extern class JspPFrameResult {
    public var data_pnt : RawPointer<Int32>;
    public var significant_changes : Int32;
}

@:include("jsplayer_cuda.h")
@:buildXml("<include name=\"${haxelib:jsplayer_cuda}/haxe/Build.xml\"/>")
extern class Jsp {
    @:native("jsp_device_count")    static function deviceCount() : Int;
    @:native("jsp_last_error")      static function lastError() : cpp.ConstCharStar;
    // new MSVideo1_16bit / MSVideo1_8bit / ScreenPressor (Manager.hx:105-111); codec = VideoData.hx:75-80 order
    @:native("jsp_create")          static function create(codec:Int, width:Int, height:Int, bpp:Int,
                                                          palette:RawConstPointer<UInt8>, paletteBytes:Int, device:Int) : RawPointer<JspDecNative>;
    @:native("jsp_destroy")         static function destroy(d:RawPointer<JspDecNative>) : Void;
    @:native("jsp_preinit")         static function preinit(d:RawPointer<JspDecNative>, insignificantLines:Int) : Void;          // Preinit
    @:native("jsp_previous_frame")  static function previousFrame(d:RawPointer<JspDecNative>) : RawPointer<Int32>;             // PreviousFrame
    @:native("jsp_is_key_frame")    static function isKeyFrame(d:RawPointer<JspDecNative>, data:RawConstPointer<UInt8>, len:Int) : Int;   // IsKeyFrame
    @:native("jsp_state_of")        static function stateOf(d:RawPointer<JspDecNative>) : Int;                                   // State
    @:native("jsp_decompress_i")    static function decompressI(d:RawPointer<JspDecNative>, src:RawConstPointer<UInt8>, len:Int, dst:RawPointer<Int32>) : Int;   // DecompressI
    @:native("jsp_continue_i")      static function continueI(d:RawPointer<JspDecNative>) : Int;                                 // ContinueI
    @:native("jsp_decompress_p")    static function decompressP(d:RawPointer<JspDecNative>, src:RawConstPointer<UInt8>, len:Int, dst:RawPointer<Int32>) : JspPFrameResult; // DecompressP
    @:native("jsp_needs_index")     static function needsIndex(d:RawPointer<JspDecNative>) : Int;                                // NeedsIndex
    @:native("jsp_stop_and_clean")  static function stopAndClean(d:RawPointer<JspDecNative>) : Void;                             // StopAndClean
}
