"""jsplayer_b200 -- B200-native (sm_100a) batch decoder for thedeemon/jsplayer's codec hot path.

The product is libjsplayer_cuda.so (CUDA kernels behind the C ABI of include/jsplayer_cuda.h).
This package is the Python host side: a ctypes mirror of the reference's IVideoCodec interface
(codec.py), the batch driver (batch.py) and the synthetic bitstream encoders (synth/).
There is no CPU decode path: every decode call raises if the CUDA library or a GPU is missing.
"""
from .codec import (DecoderState, PFrameResult, IVideoCodec, MSVideo1_16bit, MSVideo1_8bit,
                    ScreenPressor, CodecType)
from .batch import BatchDecoder, StreamSpec

__all__ = ["DecoderState", "PFrameResult", "IVideoCodec", "MSVideo1_16bit", "MSVideo1_8bit",
           "ScreenPressor", "CodecType", "BatchDecoder", "StreamSpec"]
