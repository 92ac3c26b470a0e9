"""ctypes binding of the synthetic H.264 stream generator (tools/h264gen).  Test/bench tooling."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.normpath(os.path.join(_HERE, '..', '..', 'build', 'libh264gen.so'))


class GenParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        'width', 'height', 'frames', 'gop', 'profile', 'cabac', 'bframes', 'num_ref', 'qp', 'slices')] + \
        [('seed', ctypes.c_uint32)] + [(n, ctypes.c_int32) for n in (
            'weighted', 'direct_spatial', 'deblock', 'constrained_intra', 'ipcm_per_100k', 'intra_in_p_pct',
            'cabac_init_idc', 'chroma_qp_offset', 'scaling_lists', 'poc_type', 'fragmented', 'threads',
            'qp_jitter', 'b_pyramid', 'rplm_pct', 'mmco', 'pad_refs', 'mixed_slices', 'header_variant', 'direct_4x4')] + [('reserved', ctypes.c_int32 * 1)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError('%s missing: run `python -c "import __graft_entry__ as g; g.build()"`' % _LIB_PATH)
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.hwgen_encode.argtypes = [ctypes.POINTER(GenParams), ctypes.POINTER(ctypes.c_void_p),
                                      ctypes.POINTER(ctypes.c_size_t), ctypes.c_void_p]
        _lib.hwgen_last_error.restype = ctypes.c_char_p
        _lib.hwgen_free.argtypes = [ctypes.c_void_p]
    return _lib


def generate(want_recon=False, **kw):
    """Returns mp4 bytes (and the encoder's own reconstruction, frames x (H*W*3/2) uint8, display order)."""
    L = lib()
    p = GenParams()
    L.hwgen_default_params(ctypes.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    out = ctypes.c_void_p()
    size = ctypes.c_size_t()
    recon = None
    rp = None
    if want_recon:
        recon = np.empty((p.frames, p.width * p.height * 3 // 2), np.uint8)
        rp = recon.ctypes.data
    rc = L.hwgen_encode(ctypes.byref(p), ctypes.byref(out), ctypes.byref(size), rp)
    if rc != 0:
        raise RuntimeError('hwgen_encode: ' + L.hwgen_last_error().decode())
    data = ctypes.string_at(out.value, size.value)
    L.hwgen_free(out)
    return (data, recon) if want_recon else data
