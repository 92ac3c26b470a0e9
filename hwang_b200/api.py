import ctypes
import os

import numpy as np

from . import _lib
from ._lib import EncodedDataC, Stats, c_u64p


class DeviceType:  # hwang/common.h:20-23
    CPU = 0
    GPU = 1


class VideoDecoderType:  # hwang/video_decoder_factory.h:23-27 (+ B200)
    SOFTWARE = 0
    NVIDIA = 1
    INTEL = 2
    B200 = 3


class DeviceHandle:  # hwang/common.h:25-50
    def __init__(self, type=DeviceType.CPU, id=0):
        self.type = type
        self.id = id


def device_count():
    return _lib.lib().hwb_device_count()


def _u64(seq):
    a = np.ascontiguousarray(np.asarray(seq, dtype=np.uint64))
    return a, a.ctypes.data_as(c_u64p)


class VideoIndex:
    """hwang::VideoIndex (hwang/video_index.h:22-77; python/hwang/video_index.py:5-15)."""

    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        try:
            if self._h:
                _lib.lib().hwb_video_index_destroy(self._h)
        except Exception:
            pass

    @staticmethod
    def create(timescale, duration, width, height, format, sample_offsets, sample_sizes, keyframe_indices, metadata):
        L = _lib.lib()
        so, sop = _u64(sample_offsets)
        ss, ssp = _u64(sample_sizes)
        kf, kfp = _u64(keyframe_indices)
        md = bytes(metadata)
        return VideoIndex(L.hwb_video_index_create(timescale, duration, width, height, format.encode(), sop, ssp, len(so),
                                                   kfp, len(kf), md, len(md)))

    @staticmethod
    def deserialize(data):
        data = bytes(data)
        return VideoIndex(_lib.lib().hwb_video_index_deserialize(data, len(data)))

    def serialize(self):
        L = _lib.lib()
        n = L.hwb_video_index_serialize(self._h, None, 0)
        buf = ctypes.create_string_buffer(n)
        L.hwb_video_index_serialize(self._h, buf, n)
        return buf.raw

    @staticmethod
    def from_file(path):
        with open(path, 'rb') as f:
            return VideoIndex.deserialize(f.read())

    def to_file(self, path):
        with open(path, 'wb') as f:
            f.write(self.serialize())

    def timescale(self):
        return _lib.lib().hwb_video_index_timescale(self._h)

    def duration(self):
        return _lib.lib().hwb_video_index_duration(self._h)

    def fps(self):
        return _lib.lib().hwb_video_index_fps(self._h)

    def frame_width(self):
        return _lib.lib().hwb_video_index_frame_width(self._h)

    def frame_height(self):
        return _lib.lib().hwb_video_index_frame_height(self._h)

    def format(self):
        return _lib.lib().hwb_video_index_format(self._h).decode()

    def frames(self):
        return _lib.lib().hwb_video_index_frames(self._h)

    def _arr(self, fn, n):
        p = fn(self._h)
        return [int(p[i]) for i in range(n)] if n else []

    def sample_offsets(self):
        return self._arr(_lib.lib().hwb_video_index_sample_offsets, self.frames())

    def sample_sizes(self):
        return self._arr(_lib.lib().hwb_video_index_sample_sizes, self.frames())

    def keyframe_indices(self):
        n = ctypes.c_size_t()
        p = _lib.lib().hwb_video_index_keyframe_indices(self._h, ctypes.byref(n))
        return [int(p[i]) for i in range(n.value)]

    def metadata_bytes(self):
        n = ctypes.c_size_t()
        p = _lib.lib().hwb_video_index_metadata_bytes(self._h, ctypes.byref(n))
        return ctypes.string_at(p, n.value) if n.value else b''


class MP4IndexCreator:
    """hwang::MP4IndexCreator (hwang/mp4_index_creator.h:23-45)."""

    def __init__(self, file_size):
        self._h = _lib.lib().hwb_index_creator_create(file_size)

    def __del__(self):
        try:
            _lib.lib().hwb_index_creator_destroy(self._h)
        except Exception:
            pass

    def feed(self, data, size):
        no, ns = ctypes.c_uint64(), ctypes.c_uint64()
        data = bytes(data)
        r = _lib.lib().hwb_index_creator_feed(self._h, data, min(size, len(data)), ctypes.byref(no), ctypes.byref(ns))
        return bool(r), no.value, ns.value

    def is_done(self):
        return bool(_lib.lib().hwb_index_creator_is_done(self._h))

    def is_error(self):
        return bool(_lib.lib().hwb_index_creator_is_error(self._h))

    def error_message(self):
        return _lib.lib().hwb_index_creator_error_message(self._h).decode()

    def get_video_index(self):
        return VideoIndex(_lib.lib().hwb_index_creator_get_video_index(self._h))


def slice_into_video_intervals(index, rows):
    """hwang/video_index.cpp:62-109 -> [((start, end), [rows...]), ...] like hwang_python.cpp:38-50."""
    rows = list(rows)
    if not rows:
        return []
    L = _lib.lib()
    r, rp = _u64(rows)
    cap = len(index.keyframe_indices()) + 1
    starts, ends, counts = (np.zeros(cap, np.uint64) for _ in range(3))
    valid = np.zeros(len(rows), np.uint64)
    n = L.hwb_slice_into_video_intervals(index._h, rp, len(r), starts.ctypes.data_as(c_u64p), ends.ctypes.data_as(c_u64p),
                                         counts.ctypes.data_as(c_u64p), cap, valid.ctypes.data_as(c_u64p))
    if n < 0:
        raise RuntimeError('slice_into_video_intervals: rows must be ascending frame numbers inside the video')
    out, k = [], 0
    for i in range(n):
        c = int(counts[i])
        out.append(((int(starts[i]), int(ends[i])), [int(x) for x in valid[k:k + c]]))
        k += c
    return out


class EncodedData:  # DecoderAutomata::EncodedData, hwang/decoder_automata.h:43-66
    def __init__(self):
        self.encoded_video = b''
        self.width = 0
        self.height = 0
        self.start_keyframe = 0
        self.end_keyframe = 0
        self.format = ''
        self.sample_offsets = []
        self.sample_sizes = []
        self.keyframes = []
        self.valid_frames = []


class PinnedBuffer:
    """Page-locked host memory exposed as a numpy array; output frames are views into it."""

    def __init__(self, nbytes):
        self.nbytes = nbytes
        self.ptr = _lib.lib().hwb_alloc_pinned(nbytes)
        if not self.ptr:
            raise MemoryError('hwb_alloc_pinned(%d) failed' % nbytes)
        self.array = np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(self.ptr))

    def __del__(self):
        try:
            if self.ptr:
                _lib.lib().hwb_free_pinned(self.ptr)
                self.ptr = None
        except Exception:
            pass


class DecoderAutomata:
    """hwang::DecoderAutomata (hwang/decoder_automata.h:33-70; binding hwang_python.cpp:52-98, 146-160)."""

    def __init__(self, device_handle, num_devices, decoder_type):
        self._h = _lib.lib().hwb_automata_create(device_handle.type, device_handle.id, num_devices, decoder_type)
        if not self._h:
            raise RuntimeError('DecoderAutomata: could not create decoder type %d on device (%d, %d): no such backend / '
                               'no CUDA device (there is no CPU fallback)' % (decoder_type, device_handle.type, device_handle.id))
        self._keep = None

    def __del__(self):
        try:
            if self._h:
                _lib.lib().hwb_automata_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def initialize(self, encoded_data, extradata):
        L = _lib.lib()
        arr = (EncodedDataC * len(encoded_data))()
        keep = []
        for i, d in enumerate(encoded_data):
            ev = bytes(d.encoded_video)
            so, sop = _u64(d.sample_offsets)
            ss, ssp = _u64(d.sample_sizes)
            kf, kfp = _u64(d.keyframes)
            vf, vfp = _u64(d.valid_frames)
            fmt = d.format.encode()
            keep += [ev, so, ss, kf, vf, fmt]
            arr[i].encoded_video = ctypes.cast(ctypes.c_char_p(ev), ctypes.c_void_p)
            arr[i].encoded_video_size = len(ev)
            arr[i].width, arr[i].height = d.width, d.height
            arr[i].start_keyframe, arr[i].end_keyframe = d.start_keyframe, d.end_keyframe
            arr[i].format = fmt
            arr[i].sample_offsets, arr[i].sample_sizes, arr[i].num_samples = sop, ssp, min(len(so), len(ss))
            arr[i].keyframes, arr[i].num_keyframes = kfp, len(kf)
            arr[i].valid_frames, arr[i].num_valid_frames = vfp, len(vf)
        extradata = bytes(extradata)
        if L.hwb_automata_initialize(self._h, arr, len(encoded_data), extradata, len(extradata)) != 0:
            raise RuntimeError(L.hwb_automata_last_error(self._h).decode())

    def get_frames(self, index, num_frames, pinned=True):
        """-> list of (H, W, 3) uint8 arrays.  Unlike the reference wrapper (hwang_python.cpp:72-98: one pageable
        vector + a malloc/memcpy per frame) the frames are views into one page-locked buffer the GPU copies into."""
        L = _lib.lib()
        w, h = index.frame_width(), index.frame_height()
        fs = w * h * 3
        if num_frames == 0:
            return []
        if pinned and _lib.library_path() == _lib.PRODUCT_LIB:
            buf = PinnedBuffer(fs * num_frames)
            arr, ptr = buf.array, buf.ptr
        else:
            buf = None
            arr = np.empty(fs * num_frames, np.uint8)
            ptr = arr.ctypes.data
        if L.hwb_automata_get_frames(self._h, ptr, num_frames) != 0:
            raise RuntimeError(L.hwb_automata_last_error(self._h).decode())
        frames = arr.reshape(num_frames, h, w, 3)
        out = [frames[i] for i in range(num_frames)]
        if buf is not None:
            for f in out:
                f.flags.writeable = True
            self._keep = buf  # views keep `arr` alive; `buf` owns the pinned allocation
            out = [_Owned(f, buf) for f in out]
        return out

    def stats(self):
        s = Stats()
        _lib.lib().hwb_automata_get_stats(self._h, ctypes.byref(s))
        return s.as_dict()


def _Owned(view, owner):
    """numpy view that keeps the pinned allocation alive for as long as the frame is referenced."""
    class _A(np.ndarray):
        pass
    a = view.view(_A)
    a._owner = owner
    return a


class VideoDecoder:
    """Direct access to the plugin boundary (hwang::VideoDecoderInterface, hwang/video_decoder_interface.h:26-49)."""

    def __init__(self, device_id=0, decoder_type=VideoDecoderType.B200, device_type=DeviceType.GPU, num_devices=1):
        h = ctypes.c_void_p()
        if _lib.lib().hwb_decoder_create(device_type, device_id, num_devices, decoder_type, ctypes.byref(h)) != 0 or not h:
            raise RuntimeError('VideoDecoderFactory: cannot create decoder type %d on device (%d, %d): no CUDA device / '
                               'unsupported backend (there is no CPU fallback)' % (decoder_type, device_type, device_id))
        self._h = h
        self.width = self.height = 0

    def __del__(self):
        try:
            if self._h:
                _lib.lib().hwb_decoder_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(_lib.lib().hwb_decoder_last_error(self._h).decode())

    def configure(self, width, height, format, extradata):
        extradata = bytes(extradata)
        self.width, self.height = width, height
        self._chk(_lib.lib().hwb_decoder_configure(self._h, width, height, format.encode(), extradata, len(extradata)))

    def feed(self, data, keyframe=False):
        if data is None:
            self._chk(_lib.lib().hwb_decoder_feed(self._h, None, 0, 0))
        else:
            data = bytes(data)
            self._chk(_lib.lib().hwb_decoder_feed(self._h, data, len(data), int(keyframe)))

    def flush(self):
        self._chk(_lib.lib().hwb_decoder_flush(self._h))

    def discard_frame(self):
        self._chk(_lib.lib().hwb_decoder_discard_frame(self._h))

    def decoded_frames_buffered(self):
        return _lib.lib().hwb_decoder_decoded_frames_buffered(self._h)

    def frames_ready(self):
        return _lib.lib().hwb_decoder_frames_ready(self._h)

    def wait_until_frames_copied(self):
        self._chk(_lib.lib().hwb_decoder_wait_until_frames_copied(self._h))

    def get_frame(self):
        out = np.empty((self.height, self.width, 3), np.uint8)
        self._chk(_lib.lib().hwb_decoder_get_frame(self._h, out.ctypes.data, out.nbytes))
        self.wait_until_frames_copied()
        return out

    def get_frame_into(self, ptr, nbytes):
        self._chk(_lib.lib().hwb_decoder_get_frame(self._h, ptr, nbytes))

    def get_frame_yuv(self):
        out = np.empty(self.width * self.height * 3 // 2, np.uint8)
        self._chk(_lib.lib().hwb_decoder_get_frame_yuv(self._h, out.ctypes.data, out.nbytes))
        self.wait_until_frames_copied()
        return out

    def get_frame_device(self):
        p = ctypes.c_void_p()
        self._chk(_lib.lib().hwb_decoder_get_frame_device(self._h, ctypes.byref(p)))
        return p.value

    def set_chunk_pictures(self, n):
        _lib.lib().hwb_decoder_set_chunk_pictures(self._h, n)

    def stats(self):
        s = Stats()
        _lib.lib().hwb_decoder_get_stats(self._h, ctypes.byref(s))
        return s.as_dict()


def index_video(f_or_string):
    """python/hwang/__init__.py:5-26: pull-parse an MP4 into a VideoIndex."""
    def w(f):
        f.seek(0, os.SEEK_END)
        size = f.tell()
        f.seek(0, 0)
        indexer = MP4IndexCreator(size)
        offset = 0
        size_to_read = 1024
        while not indexer.is_done():
            f.seek(offset, 0)
            data = f.read(size_to_read)
            ret, offset, new_size = indexer.feed(data, size_to_read)
            size_to_read = new_size
        if indexer.is_error():
            raise Exception(indexer.error_message())
        return indexer.get_video_index()

    if isinstance(f_or_string, str):
        with open(f_or_string, 'rb') as f:
            return w(f)
    return w(f_or_string)


class Decoder(object):
    """python/hwang/decoder.py:5-69.  device_type defaults to GPU here (the only backend); CPU raises."""

    def __init__(self, f_or_path, video_index=None, device_type=DeviceType.GPU, device_id=0):
        if video_index is None:
            video_index = index_video(f_or_path)
        self.video_index = video_index
        self.f = open(f_or_path, 'rb') if isinstance(f_or_path, str) else f_or_path
        handle = DeviceHandle(device_type, device_id)
        decoder_type = VideoDecoderType.SOFTWARE
        if device_type == DeviceType.GPU:
            decoder_type = VideoDecoderType.NVIDIA  # the reference's mapping (decoder.py:25-28); the factory routes it to B200
        self._decoder = DecoderAutomata(handle, 1, decoder_type)

    def retrieve(self, rows):
        video_intervals = slice_into_video_intervals(self.video_index, rows)
        frames = []
        sample_offsets = self.video_index.sample_offsets()
        sample_sizes = self.video_index.sample_sizes()
        sample_offsets.append(sample_offsets[-1] + sample_sizes[-1])
        sample_sizes.append(0)
        keyframe_indices = self.video_index.keyframe_indices()
        for (start_index, end_index), valid_frames in video_intervals:
            start_offset = sample_offsets[start_index]
            end_offset = sample_offsets[end_index] + sample_sizes[end_index]
            self.f.seek(start_offset, 0)
            encoded_data = self.f.read(end_offset - start_offset)
            data = EncodedData()
            data.width = self.video_index.frame_width()
            data.height = self.video_index.frame_height()
            data.format = self.video_index.format()
            data.start_keyframe = start_index
            data.end_keyframe = end_index
            data.sample_offsets = [o - start_offset for o in sample_offsets[start_index:end_index]]
            data.sample_sizes = sample_sizes[start_index:end_index]
            data.valid_frames = valid_frames
            data.keyframes = [k for k in keyframe_indices if k >= start_index and k <= end_index]
            data.encoded_video = encoded_data
            self._decoder.initialize([data], self.video_index.metadata_bytes())
            frames += self._decoder.get_frames(self.video_index, len(valid_frames))
        return frames
