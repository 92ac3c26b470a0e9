import ctypes
import os
import threading

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


# Page-locked allocations are slow (cudaHostAlloc pins page by page: a few GB/s), and a sparse request returns a gigabyte
# of frames: released buffers are kept for the next request instead of being unpinned.  HWB_PINNED_POOL_MB bounds what
# the pool may hold (default 6144; 0 = no pool).
_pool_lock = threading.Lock()
_pool = []  # (capacity, ptr), idle buffers
_pool_bytes = 0
_POOL_MAX = int(os.environ.get('HWB_PINNED_POOL_MB', '6144')) << 20
_POOL_GRAIN = 8 << 20


def _pool_take(nbytes):
    global _pool_bytes
    with _pool_lock:
        best = None
        for i, (cap, ptr) in enumerate(_pool):
            if cap >= nbytes and cap <= 2 * nbytes + _POOL_GRAIN and (best is None or cap < _pool[best][0]):
                best = i
        if best is None:
            return None
        cap, ptr = _pool.pop(best)
        _pool_bytes -= cap
        return cap, ptr


def _pool_give(cap, ptr):
    """-> list of pointers the caller has to free (the buffer itself if the pool is off, evicted ones otherwise)."""
    global _pool_bytes
    if _POOL_MAX <= 0 or cap > _POOL_MAX:
        return [ptr]
    evict = []
    with _pool_lock:
        _pool.append((cap, ptr))
        _pool_bytes += cap
        while _pool_bytes > _POOL_MAX:  # oldest first
            c, p = _pool.pop(0)
            _pool_bytes -= c
            evict.append(p)
    return evict


def release_pinned_pool():
    """Unpin every idle buffer of the pool."""
    global _pool_bytes
    with _pool_lock:
        idle, _pool[:] = list(_pool), []
        _pool_bytes = 0
    for _, p in idle:
        _lib.lib().hwb_free_pinned(p)


class PinnedBuffer:
    """Page-locked host memory exposed as a numpy array; output frames are views into it."""

    def __init__(self, nbytes):
        self.nbytes = nbytes
        got = _pool_take(nbytes) if _POOL_MAX > 0 else None
        if got:
            self.capacity, self.ptr = got
        else:
            self.capacity = (nbytes + _POOL_GRAIN - 1) // _POOL_GRAIN * _POOL_GRAIN
            self.ptr = _lib.lib().hwb_alloc_pinned(self.capacity)
            if not self.ptr:
                release_pinned_pool()  # the idle buffers may be what is in the way
                self.ptr = _lib.lib().hwb_alloc_pinned(self.capacity)
            if not self.ptr:
                raise MemoryError('hwb_alloc_pinned(%d) failed' % self.capacity)
        self.array = np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(self.ptr))

    def __del__(self):
        try:
            if self.ptr:
                ptr, self.ptr = self.ptr, None
                for p in _pool_give(self.capacity, ptr):
                    _lib.lib().hwb_free_pinned(p)
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

    def get_frames_into(self, ptr, num_frames):
        """num_frames tightly packed RGB24 frames to caller-owned memory at `ptr`: pageable or page-locked host memory,
        or device memory (DeviceType::GPU output)."""
        L = _lib.lib()
        if L.hwb_automata_get_frames(self._h, ptr, num_frames) != 0:
            raise RuntimeError(L.hwb_automata_last_error(self._h).decode())

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

    def set_interval_hint(self, start_frame, wanted):
        """Before the first feed() of an interval: the frames that will be fetched (the rest will be discarded)."""
        w, wp = _u64(wanted)
        _lib.lib().hwb_decoder_set_interval_hint(self._h, start_frame, wp, len(w))

    def set_defer_submit(self, on):
        """Collect the pictures of consecutive intervals / clips of equal geometry into one GPU batch (see hwang_b200.h)."""
        self._chk(_lib.lib().hwb_decoder_set_defer_submit(self._h, int(bool(on))))

    def submit_pending(self):
        self._chk(_lib.lib().hwb_decoder_submit_pending(self._h))

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


class DeviceFrames:
    """n RGB24 frames (n, H, W, 3) uint8 in DEVICE memory (hwang::DeviceType::GPU as the output location,
    hwang/common.h:20-50).  Implements __cuda_array_interface__, so `torch.as_tensor(frames, device='cuda:%d' % id)`
    / cupy.asarray wrap it without a copy; to_host() brings it back for callers without a CUDA runtime."""

    def __init__(self, device_id, n, height, width):
        self.device_id, self.shape = device_id, (n, height, width, 3)
        self.nbytes = n * height * width * 3
        self.ptr = _lib.lib().hwb_alloc_device(device_id, max(1, self.nbytes))
        if not self.ptr:
            raise MemoryError('hwb_alloc_device(%d bytes) failed' % self.nbytes)

    @property
    def __cuda_array_interface__(self):
        return {'shape': self.shape, 'typestr': '|u1', 'data': (self.ptr, False), 'version': 2, 'strides': None}

    def to_host(self):
        out = np.empty(self.shape, np.uint8)
        if self.nbytes and _lib.lib().hwb_copy_device_to_host(self.device_id, out.ctypes.data, self.ptr, self.nbytes) != 0:
            raise RuntimeError('device to host copy failed')
        return out

    def __len__(self):
        return self.shape[0]

    def __del__(self):
        try:
            if self.ptr:
                _lib.lib().hwb_free_device(self.device_id, self.ptr)
                self.ptr = None
        except Exception:
            pass


def encoded_intervals(f, video_index, rows):
    """rows -> one EncodedData per keyframe-delimited interval that holds wanted rows (what the body of the
    reference's Decoder.retrieve loop builds, python/hwang/decoder.py:41-64: one contiguous read per interval, offsets
    rebased to it)."""
    offs, sizes = video_index.sample_offsets(), video_index.sample_sizes()
    kfs = video_index.keyframe_indices()
    w, h, fmt = video_index.frame_width(), video_index.frame_height(), video_index.format()
    out = []
    for (a, b), valid in slice_into_video_intervals(video_index, rows):
        ed = EncodedData()
        ed.width, ed.height, ed.format = w, h, fmt
        ed.start_keyframe, ed.end_keyframe = a, b
        base = offs[a]
        f.seek(base, 0)
        ed.encoded_video = f.read(offs[b - 1] + sizes[b - 1] - base)
        ed.sample_offsets = [o - base for o in offs[a:b]]
        ed.sample_sizes = sizes[a:b]
        ed.valid_frames = list(valid)
        ed.keyframes = [k for k in kfs if a <= k <= b]
        out.append(ed)
    return out


class Decoder(object):
    """hwang.Decoder (python/hwang/decoder.py:5-69): sparse random-access frame retrieval from one video.
    device_type defaults to GPU here (the only backend); CPU raises.

    Unlike the reference, which initialises the automaton once per interval (decoder.py:65), a request's intervals go
    to the automaton in ONE initialize call: its feeder collects their pictures into common GPU batches, so the
    (latency-bound) entropy stage sees the slices of all intervals at once."""

    def __init__(self, f_or_path, video_index=None, device_type=DeviceType.GPU, device_id=0):
        if video_index is None:
            video_index = index_video(f_or_path)
        self.video_index = video_index
        self.f = open(f_or_path, 'rb') if isinstance(f_or_path, str) else f_or_path
        self.device_id = device_id
        # the reference's mapping (decoder.py:25-28): GPU -> NVIDIA, which the factory routes to the B200 backend
        decoder_type = VideoDecoderType.NVIDIA if device_type == DeviceType.GPU else VideoDecoderType.SOFTWARE
        self._decoder = DecoderAutomata(DeviceHandle(device_type, device_id), 1, decoder_type)

    def _start(self, rows):
        intervals = encoded_intervals(self.f, self.video_index, rows)
        if intervals:
            self._decoder.initialize(intervals, self.video_index.metadata_bytes())
        return sum(len(ed.valid_frames) for ed in intervals)

    def retrieve(self, rows):
        """rows (ascending frame numbers) -> list of (H, W, 3) uint8 arrays (views of one page-locked buffer)."""
        n = self._start(rows)
        return self._decoder.get_frames(self.video_index, n) if n else []

    def retrieve_device(self, rows):
        """As retrieve(), but the frames stay in device memory: -> DeviceFrames (n, H, W, 3)."""
        n = self._start(rows)
        out = DeviceFrames(self.device_id, n, self.video_index.frame_height(), self.video_index.frame_width())
        if n:
            self._decoder.get_frames_into(out.ptr, n)
        return out
