"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of the reference's software decode path,
`hwang/impls/software/software_video_decoder.cpp:103-457`, against the libavcodec / libswscale
shared objects that ship inside the image's opencv wheel (FFmpeg 8.0; the reference pins FFmpeg
n3.3.1, `deps.sh:148`, which cannot be built here -- H.264 decoding is normative/bit-exact so the
decoder version does not matter on conformant streams; swscale parity is pinned separately by
`yuv420_to_rgb24`).

Call sequence restated (file:line in the reference):
  configure  :103-165  -> `FFmpegH264.__init__` (avcodec_find_decoder / alloc_context3 / thread_count / open2)
  feed       :167-248  -> `avcc_to_annexb` (what the h264_mp4toannexb BSF does: length prefixes ->
                          start codes, SPS/PPS from avcC prepended at keyframes) + `send`
  feed_packet:349-402  -> `send` / `_drain` (avcodec_send_packet, avcodec_receive_frame loop)
  flush      :250-268  -> `flush`
  get_frame  :281-339  -> `sws_rgb24` (sws_getContext(.., RGB24, SWS_BICUBIC) + sws_scale, stride W*3)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this.
"""
import ctypes
import glob
import os
import struct

import numpy as np

_LIBS = None
AVERROR_EAGAIN = -11
AVERROR_EOF = -541478725
AV_CODEC_ID_H264 = 27


def _load():
    global _LIBS
    if _LIBS is not None:
        return _LIBS
    import cv2  # noqa: F401  (resolves the bundled libraries' own dependencies)
    base = os.path.join(os.path.dirname(cv2.__file__), '..', 'opencv_python_headless.libs')
    def lib(name):
        return ctypes.CDLL(glob.glob(os.path.join(base, name + '-*.so*'))[0])
    avutil = lib('libavutil')
    avcodec = lib('libavcodec')
    swscale = lib('libswscale')
    P = ctypes.c_void_p
    avcodec.avcodec_find_decoder.restype = P
    avcodec.avcodec_find_decoder.argtypes = [ctypes.c_int]
    avcodec.avcodec_alloc_context3.restype = P
    avcodec.avcodec_alloc_context3.argtypes = [P]
    avcodec.avcodec_open2.argtypes = [P, P, P]
    avcodec.avcodec_send_packet.argtypes = [P, P]
    avcodec.avcodec_receive_frame.argtypes = [P, P]
    avcodec.avcodec_flush_buffers.argtypes = [P]
    avcodec.avcodec_free_context.argtypes = [ctypes.POINTER(P)]
    avcodec.av_packet_alloc.restype = P
    avcodec.av_new_packet.argtypes = [P, ctypes.c_int]
    avcodec.av_packet_unref.argtypes = [P]
    avcodec.av_packet_free.argtypes = [ctypes.POINTER(P)]
    avutil.av_frame_alloc.restype = P
    avutil.av_frame_unref.argtypes = [P]
    avutil.av_frame_free.argtypes = [ctypes.POINTER(P)]
    avutil.av_opt_set_int.argtypes = [P, ctypes.c_char_p, ctypes.c_int64, ctypes.c_int]
    avutil.av_log_set_level.argtypes = [ctypes.c_int]
    swscale.sws_getContext.restype = P
    swscale.sws_getContext.argtypes = [ctypes.c_int] * 3 + [ctypes.c_int] * 3 + [ctypes.c_int, P, P, P]
    swscale.sws_scale.argtypes = [P, P, P, ctypes.c_int, ctypes.c_int, P, P]
    swscale.sws_freeContext.argtypes = [P]
    avcodec.avcodec_version.restype = ctypes.c_uint
    _LIBS = (avutil, avcodec, swscale)
    return _LIBS


def ffmpeg_version():
    _, avcodec, _ = _load()
    v = avcodec.avcodec_version()
    return 'libavcodec %d.%d.%d' % (v >> 16, (v >> 8) & 255, v & 255)


def parse_avcc(avcc):
    """avcC -> (nal_length_size, [sps...], [pps...])   (reference: software_video_decoder.cpp:155-160)"""
    avcc = bytes(avcc)
    nls = (avcc[4] & 3) + 1
    nsps = avcc[5] & 31
    off = 6
    sps, pps = [], []
    for _ in range(nsps):
        n = struct.unpack('>H', avcc[off:off + 2])[0]
        sps.append(avcc[off + 2:off + 2 + n]); off += 2 + n
    npps = avcc[off]; off += 1
    for _ in range(npps):
        n = struct.unpack('>H', avcc[off:off + 2])[0]
        pps.append(avcc[off + 2:off + 2 + n]); off += 2 + n
    return nls, sps, pps


def avcc_to_annexb(sample, nls, sps, pps, keyframe):
    """What h264_mp4toannexb + the keyframe extradata prepend do (software_video_decoder.cpp:173-202)."""
    out = bytearray()
    if keyframe:
        for n in list(sps) + list(pps):
            out += b'\x00\x00\x00\x01' + n
    off = 0
    sample = bytes(sample)
    while off + nls <= len(sample):
        n = int.from_bytes(sample[off:off + nls], 'big')
        off += nls
        out += b'\x00\x00\x00\x01' + sample[off:off + n]
        off += n
    return bytes(out)


class FFmpegH264:
    """One libavcodec h264 decoder instance; frames come out in display order as (Y, U, V) uint8 arrays."""

    def __init__(self, threads=1, quiet=True):
        avutil, avcodec, _ = _load()
        self.avutil, self.avcodec = avutil, avcodec
        if quiet:
            avutil.av_log_set_level(-8)
        codec = avcodec.avcodec_find_decoder(AV_CODEC_ID_H264)
        self.ctx = ctypes.c_void_p(avcodec.avcodec_alloc_context3(codec))
        avutil.av_opt_set_int(self.ctx, b'threads', threads, 0)
        if avcodec.avcodec_open2(self.ctx, codec, None) < 0:
            raise RuntimeError('avcodec_open2 failed')
        self.pkt = ctypes.c_void_p(avcodec.av_packet_alloc())
        self.frame = ctypes.c_void_p(avutil.av_frame_alloc())

    def close(self):
        if self.ctx:
            self.avcodec.av_packet_free(ctypes.byref(self.pkt))
            self.avutil.av_frame_free(ctypes.byref(self.frame))
            self.avcodec.avcodec_free_context(ctypes.byref(self.ctx))
            self.ctx = None

    def _read_frame(self, copy=True):
        f = self.frame.value
        data = (ctypes.c_void_p * 3).from_address(f + 0)
        ls = (ctypes.c_int * 3).from_address(f + 64)
        w = ctypes.c_int.from_address(f + 104).value
        h = ctypes.c_int.from_address(f + 108).value
        planes = []
        for i in range(3):
            pw, ph = (w, h) if i == 0 else ((w + 1) // 2, (h + 1) // 2)
            buf = (ctypes.c_uint8 * (ls[i] * ph)).from_address(data[i])
            a = np.frombuffer(buf, dtype=np.uint8).reshape(ph, ls[i])[:, :pw]
            planes.append(a.copy() if copy else a)
        return tuple(planes)

    def _drain(self, sink):
        n = 0
        while True:
            r = self.avcodec.avcodec_receive_frame(self.ctx, self.frame)
            if r == AVERROR_EAGAIN or r == AVERROR_EOF:
                return n
            if r < 0:
                raise RuntimeError('avcodec_receive_frame error %d' % r)
            if sink is not None:
                sink(self._read_frame())
            n += 1
            self.avutil.av_frame_unref(self.frame)

    def send_raw(self, annexb, on_frame):
        """As send(), but every decoded AVFrame is handed to on_frame(address of the AVFrame) before it is released:
        no plane is copied on the Python side (the benchmark's reference arm converts straight from the AVFrame, as
        SoftwareVideoDecoder::get_frame does, software_video_decoder.cpp:300-325)."""
        if annexb is None:
            r = self.avcodec.avcodec_send_packet(self.ctx, None)
        else:
            self.avcodec.av_new_packet(self.pkt, len(annexb))
            dptr = ctypes.c_void_p.from_address(self.pkt.value + 24).value
            ctypes.memmove(dptr, annexb, len(annexb))
            r = self.avcodec.avcodec_send_packet(self.ctx, self.pkt)
            self.avcodec.av_packet_unref(self.pkt)
        if r < 0 and r != AVERROR_EOF:
            raise RuntimeError('avcodec_send_packet error %d' % r)
        n = 0
        while True:
            r = self.avcodec.avcodec_receive_frame(self.ctx, self.frame)
            if r == AVERROR_EAGAIN or r == AVERROR_EOF:
                return n
            if r < 0:
                raise RuntimeError('avcodec_receive_frame error %d' % r)
            on_frame(self.frame.value)
            n += 1
            self.avutil.av_frame_unref(self.frame)

    def send(self, annexb, sink):
        """avcodec_send_packet(annexb) (None = drain signal) then receive until EAGAIN/EOF."""
        if annexb is None:
            r = self.avcodec.avcodec_send_packet(self.ctx, None)
        else:
            self.avcodec.av_new_packet(self.pkt, len(annexb))
            dptr = ctypes.c_void_p.from_address(self.pkt.value + 24).value
            ctypes.memmove(dptr, annexb, len(annexb))
            r = self.avcodec.avcodec_send_packet(self.ctx, self.pkt)
            self.avcodec.av_packet_unref(self.pkt)
        if r < 0 and r != AVERROR_EOF:
            raise RuntimeError('avcodec_send_packet error %d' % r)
        return self._drain(sink)

    def flush(self):
        self.avcodec.avcodec_flush_buffers(self.ctx)


def decode_samples(avcc, samples, keyflags, threads=1, sink=None):
    """Decode MP4 samples (decode order) the way the reference does; returns frames in display order
    (or streams them to `sink`)."""
    nls, sps, pps = parse_avcc(avcc)
    dec = FFmpegH264(threads, quiet=not bool(os.environ.get("FFLOG")))
    frames = []
    out = sink if sink is not None else frames.append
    try:
        for s, k in zip(samples, keyflags):
            dec.send(avcc_to_annexb(s, nls, sps, pps, k), out)
        dec.send(None, out)
        dec.flush()
    finally:
        dec.close()
    return frames


class SwsRgb24:
    """sws_getContext(w,h,YUV420P -> w,h,RGB24, SWS_BICUBIC) + sws_scale (software_video_decoder.cpp:292-325)."""

    def __init__(self, w, h):
        _, _, sws = _load()
        self.sws, self.w, self.h = sws, w, h
        self.ctx = ctypes.c_void_p(sws.sws_getContext(w, h, 0, w, h, 2, 4, None, None, None))

    def __call__(self, y, u, v):
        y, u, v = (np.ascontiguousarray(p) for p in (y, u, v))
        dst = np.empty((self.h, self.w, 3), np.uint8)
        src = (ctypes.c_void_p * 4)(y.ctypes.data, u.ctypes.data, v.ctypes.data, None)
        sst = (ctypes.c_int * 4)(y.strides[0], u.strides[0], v.strides[0], 0)
        dp = (ctypes.c_void_p * 4)(dst.ctypes.data, None, None, None)
        dst_st = (ctypes.c_int * 4)(self.w * 3, 0, 0, 0)
        self.sws.sws_scale(self.ctx, src, sst, 0, self.h, dp, dst_st)
        return dst

    def scale_avframe(self, frame_addr, dst_addr):
        """sws_scale straight from an AVFrame (data[] at offset 0, linesize[] at offset 64) into dst (W*3 stride):
        the call at software_video_decoder.cpp:325 with no intermediate copy."""
        dp = (ctypes.c_void_p * 4)(dst_addr, None, None, None)
        dst_st = (ctypes.c_int * 4)(self.w * 3, 0, 0, 0)
        self.sws.sws_scale(self.ctx, ctypes.c_void_p(frame_addr), ctypes.c_void_p(frame_addr + 64), 0, self.h, dp, dst_st)

    def close(self):
        if self.ctx:
            self.sws.sws_freeContext(self.ctx)
            self.ctx = None


def yuv420_to_rgb24(y, u, v):
    """numpy restatement of the arithmetic swscale's unscaled yuv420p->rgb24 path performs
    (exhaustively matched over all 2^24 (Y,U,V) triples)."""
    Y = y.astype(np.int32)
    U = np.repeat(np.repeat(u.astype(np.int32), 2, 0), 2, 1)[:Y.shape[0], :Y.shape[1]]
    V = np.repeat(np.repeat(v.astype(np.int32), 2, 0), 2, 1)[:Y.shape[0], :Y.shape[1]]
    yy = ((Y * 8 - 128) * 9539) >> 16
    uu = (U - 128) * 8
    vv = (V - 128) * 8
    r = yy + ((vv * 13075) >> 16)
    g = yy + ((uu * -3209) >> 16) + ((vv * -6660) >> 16)
    b = yy + ((uu * 16525) >> 16)
    return np.clip(np.stack([r, g, b], -1), 0, 255).astype(np.uint8)
