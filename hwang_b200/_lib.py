"""ctypes binding of the C ABI declared in include/hwang_b200.h.

The product library is hwang_b200/libhwang_b200.so (C++ host side + CUDA kernels, built in-tree by
hwang_b200/build.py).  There is no CPU fallback: if the library is missing, or no CUDA device is
present, decoder creation fails loudly.  Unit tests may point the binding at the host-emulation
build (tests/emu/libhwb_emu.so) with use_library(); nothing in the package does that on its own.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.environ.get('HWB_PRODUCT_LIB') or os.path.join(_HERE, 'libhwang_b200.so')  # override: A/B runs of two builds on one box
_lib = None
_path = None

c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_u64p = ctypes.POINTER(ctypes.c_uint64)


class Stats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in ('pictures_decoded', 'frames_returned', 'chunks', 'bitstream_bytes',
                                               'kernel_launches', 'h2d_bytes', 'd2h_bytes', 'algorithmic_bytes')] + \
               [(n, ctypes.c_double) for n in ('wall_ms', 'entropy_ms', 'picture_ms')] + \
               [(n, ctypes.c_uint64) for n in ('entropy_launches', 'picture_launches', 'aux_launches')]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class EncodedDataC(ctypes.Structure):
    _fields_ = [('encoded_video', ctypes.c_void_p), ('encoded_video_size', ctypes.c_size_t),
                ('width', ctypes.c_uint32), ('height', ctypes.c_uint32),
                ('start_keyframe', ctypes.c_uint64), ('end_keyframe', ctypes.c_uint64),
                ('format', ctypes.c_char_p),
                ('sample_offsets', c_u64p), ('sample_sizes', c_u64p), ('num_samples', ctypes.c_size_t),
                ('keyframes', c_u64p), ('num_keyframes', ctypes.c_size_t),
                ('valid_frames', c_u64p), ('num_valid_frames', ctypes.c_size_t)]


# name -> (restype, argtypes); every symbol include/hwang_b200.h declares
V, I, U32, U64, SZ, P, CP, D = (None, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_size_t, ctypes.c_void_p,
                                ctypes.c_char_p, ctypes.c_double)
PP = ctypes.POINTER(ctypes.c_void_p)
SIGNATURES = {
    'hwb_version': (CP, []),
    'hwb_device_count': (I, []),
    'hwb_has_decoder_type': (I, [I]),
    'hwb_decoder_create': (I, [I, I, I, I, PP]),
    'hwb_decoder_destroy': (V, [P]),
    'hwb_decoder_configure': (I, [P, U32, U32, CP, P, SZ]),
    'hwb_decoder_feed': (I, [P, P, SZ, I]),
    'hwb_decoder_flush': (I, [P]),
    'hwb_decoder_discard_frame': (I, [P]),
    'hwb_decoder_get_frame': (I, [P, P, SZ]),
    'hwb_decoder_decoded_frames_buffered': (I, [P]),
    'hwb_decoder_wait_until_frames_copied': (I, [P]),
    'hwb_decoder_last_error': (CP, [P]),
    'hwb_decoder_get_frame_yuv': (I, [P, P, SZ]),
    'hwb_decoder_get_frame_device': (I, [P, PP]),
    'hwb_decoder_frames_ready': (I, [P]),
    'hwb_decoder_set_chunk_pictures': (I, [P, I]),
    'hwb_decoder_set_interval_hint': (I, [P, U64, P, SZ]),
    'hwb_decoder_set_defer_submit': (I, [P, I]),
    'hwb_decoder_submit_pending': (I, [P]),
    'hwb_decoder_get_stats': (I, [P, ctypes.POINTER(Stats)]),
    'hwb_alloc_pinned': (P, [SZ]),
    'hwb_free_pinned': (V, [P]),
    'hwb_alloc_device': (P, [I, SZ]),
    'hwb_free_device': (V, [I, P]),
    'hwb_copy_device_to_host': (I, [I, P, P, SZ]),
    'hwb_index_creator_create': (P, [U64]),
    'hwb_index_creator_destroy': (V, [P]),
    'hwb_index_creator_feed': (I, [P, P, SZ, c_u64p, c_u64p]),
    'hwb_index_creator_is_done': (I, [P]),
    'hwb_index_creator_is_error': (I, [P]),
    'hwb_index_creator_error_message': (CP, [P]),
    'hwb_index_creator_get_video_index': (P, [P]),
    'hwb_video_index_create': (P, [U32, U64, U32, U32, CP, c_u64p, c_u64p, SZ, c_u64p, SZ, P, SZ]),
    'hwb_video_index_destroy': (V, [P]),
    'hwb_video_index_deserialize': (P, [P, SZ]),
    'hwb_video_index_serialize': (SZ, [P, P, SZ]),
    'hwb_video_index_timescale': (U32, [P]),
    'hwb_video_index_duration': (U64, [P]),
    'hwb_video_index_fps': (D, [P]),
    'hwb_video_index_frame_width': (U32, [P]),
    'hwb_video_index_frame_height': (U32, [P]),
    'hwb_video_index_format': (CP, [P]),
    'hwb_video_index_frames': (U64, [P]),
    'hwb_video_index_sample_offsets': (c_u64p, [P]),
    'hwb_video_index_sample_sizes': (c_u64p, [P]),
    'hwb_video_index_keyframe_indices': (c_u64p, [P, ctypes.POINTER(SZ)]),
    'hwb_video_index_metadata_bytes': (c_u8p, [P, ctypes.POINTER(SZ)]),
    'hwb_slice_into_video_intervals': (I, [P, c_u64p, SZ, c_u64p, c_u64p, c_u64p, SZ, c_u64p]),
    'hwb_automata_create': (P, [I, I, I, I]),
    'hwb_automata_destroy': (V, [P]),
    'hwb_automata_initialize': (I, [P, ctypes.POINTER(EncodedDataC), SZ, P, SZ]),
    'hwb_automata_get_frames': (I, [P, P, ctypes.c_int32]),
    'hwb_automata_set_chunk_pictures': (I, [P, I]),
    'hwb_automata_last_error': (CP, [P]),
    'hwb_automata_get_stats': (I, [P, ctypes.POINTER(Stats)]),
}


def use_library(path):
    """Bind to an explicit shared object (tests: the host-emulation build)."""
    global _lib, _path
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib, _path = lib, path
    return lib


def lib():
    if _lib is None:
        if not os.path.exists(PRODUCT_LIB):
            raise RuntimeError('hwang_b200: %s is missing -- build it with `python hwang_b200/build.py` '
                               '(there is no CPU fallback)' % PRODUCT_LIB)
        use_library(PRODUCT_LIB)
    return _lib


def library_path():
    return _path
