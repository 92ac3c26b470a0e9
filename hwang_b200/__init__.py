"""hwang_b200: B200-native H.264 frame-decode engine behind hwang's Python API.

Same public names as the reference package (python/hwang/__init__.py:5-26, decoder.py:5-69,
video_index.py:5-15 and the pybind module hwang/hwang_python.cpp:102-167):
  index_video, Decoder(...).retrieve(rows), VideoIndex(.from_file/.to_file), MP4IndexCreator,
  DeviceType, DeviceHandle, VideoDecoderType, EncodedData, DecoderAutomata, slice_into_video_intervals.
"""
from .api import (DeviceType, DeviceHandle, VideoDecoderType, VideoIndex, MP4IndexCreator, EncodedData,  # noqa: F401
                  DecoderAutomata, VideoDecoder, slice_into_video_intervals, index_video, Decoder, device_count)
