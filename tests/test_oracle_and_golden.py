"""CPU tier: pins the oracle (and the stream generator) against the committed golden fixtures and against the
real libraries, then checks the decode core (host emulation of the device code) against both."""
import glob
import hashlib
import io
import json
import os

import numpy as np
import pytest

import hwang_b200 as hw
from hwang_b200.testing import streamgen
from oracle import ffmpeg_oracle as fo, intervals as oiv, mp4_simple
import hwb_testutil as util

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
NAMES = sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(GOLDEN, '*.json')))


def load(name):
    with open(os.path.join(GOLDEN, name + '.mp4'), 'rb') as f:
        mp4 = f.read()
    with open(os.path.join(GOLDEN, name + '.json')) as f:
        return mp4, json.load(f)


def split(mp4):
    idx = mp4_simple.index_mp4(mp4)
    kf = set(idx['keyframes'])
    samples = [mp4[o:o + s] for o, s in zip(idx['offsets'], idx['sizes'])]
    return idx, samples, [i in kf for i in range(len(samples))]


@pytest.mark.parametrize('name', NAMES)
def test_oracle_reproduces_golden_yuv_and_rgb(name):
    """libavcodec + libswscale in this image reproduce the committed MD5s (decoder and swscale are deterministic)."""
    mp4, g = load(name)
    idx, samples, kf = split(mp4)
    frames = fo.decode_samples(idx['avcc'], samples, kf)
    sws = fo.SwsRgb24(idx['width'], idx['height'])
    assert len(frames) == len(g['yuv_md5'])
    for i, (y, u, v) in enumerate(frames):
        assert hashlib.md5(y.tobytes() + u.tobytes() + v.tobytes()).hexdigest() == g['yuv_md5'][i]
        assert hashlib.md5(sws(y, u, v).tobytes()).hexdigest() == g['rgb_md5'][i]
        # the numpy restatement of the swscale arithmetic is bit-exact against the golden RGB
        assert hashlib.md5(fo.yuv420_to_rgb24(y, u, v).tobytes()).hexdigest() == g['rgb_md5'][i]


@pytest.mark.parametrize('name', NAMES)
def test_generator_is_deterministic(built, name):
    mp4, g = load(name)
    assert streamgen.generate(**g['params']) == mp4


@pytest.mark.parametrize('name', NAMES)
def test_generator_reconstruction_equals_libavcodec(built, name):
    """The generator is a closed-loop encoder built on the decode core's own prediction / transform / deblock code:
    its reconstruction must be what libavcodec decodes -- this pins that code without any entropy decoder."""
    mp4, g = load(name)
    _, recon = streamgen.generate(want_recon=True, **g['params'])
    for i, md5 in enumerate(g['yuv_md5']):
        assert hashlib.md5(recon[i].tobytes()).hexdigest() == md5


@pytest.mark.parametrize('name', NAMES)
def test_decode_core_emulation_matches_golden(emu, name):
    mp4, g = load(name)
    index = hw.index_video(io.BytesIO(mp4))
    offs, sizes, kfs = index.sample_offsets(), index.sample_sizes(), set(index.keyframe_indices())
    samples = [mp4[o:o + s] for o, s in zip(offs, sizes)]
    got, dec = util.decode_yuv(index, samples, [i in kfs for i in range(len(samples))], chunk_pictures=1)
    assert [hashlib.md5(f.tobytes()).hexdigest() for f in got] == g['yuv_md5']
    # RGB through the automaton (get_frame path)
    rows = list(range(len(samples)))
    frames = hw.Decoder(io.BytesIO(mp4), video_index=index).retrieve(rows)
    assert [hashlib.md5(np.asarray(f).tobytes()).hexdigest() for f in frames] == g['rgb_md5']


@pytest.mark.parametrize('name', sorted(k for k in util.FEATURE_CLIPS if k != 'cropped_1080'))
def test_decode_core_emulation_feature_clips(emu, name):
    kw = dict(util.FEATURE_CLIPS[name])
    kw['frames'] = min(kw['frames'], 12)
    kw['gop'] = min(kw['gop'], 6)
    util.assert_yuv_parity(kw)


@pytest.mark.parametrize('name', sorted(util.SYNTAX_CLIPS) + sorted(util.CPU_SYNTAX_CLIPS))
def test_real_encoder_syntax_generator_and_decode_core_match_libavcodec(emu, name):
    """Reference B pictures, ref_pic_list_modification, MMCO 1-6 / long-term references, POC types 1 and 2 with MMCO 5:
    the generator's own reconstruction equals libavcodec's output (so the stream means what the generator thinks it
    means) and the decode core (host parser + emulated device) equals libavcodec frame by frame, in display order."""
    kw = util.SYNTAX_CLIPS.get(name) or util.CPU_SYNTAX_CLIPS[name]
    mp4, recon = streamgen.generate(want_recon=True, **kw)
    index = hw.index_video(io.BytesIO(mp4))
    offs, sizes, kfs = index.sample_offsets(), index.sample_sizes(), set(index.keyframe_indices())
    samples = [mp4[o:o + s] for o, s in zip(offs, sizes)]
    keyflags = [i in kfs for i in range(len(samples))]
    ref = util.oracle_frames(index, samples, keyflags)
    assert len(ref) == kw['frames']
    for i, r in enumerate(ref):
        assert np.array_equal(recon[i], util.flat(r)), 'generator reconstruction differs from libavcodec at frame %d' % i
    got, _ = util.decode_yuv(index, samples, keyflags)
    for i, (g, r) in enumerate(zip(got, ref)):
        assert np.array_equal(g, util.flat(r)), 'frame %d differs from libavcodec' % i


@pytest.mark.parametrize('name', NAMES)
def test_intervals_restatement_pinned_by_reference(name):
    mp4, g = load(name)
    ri = g['reference_index']
    for case in g['reference_intervals']:
        got = oiv.slice_into_video_intervals(ri['offsets'], ri['sizes'], ri['keyframes'], ri['frames'], case['rows'])
        exp = [((d['start'], d['end']), d['rows']) for d in case['intervals']]
        assert got == exp


def test_rgb_restatement_matches_swscale_on_random_planes():
    rng = np.random.default_rng(0)
    for (w, h) in ((64, 48), (640, 480), (1920, 1080)):
        y = rng.integers(0, 256, (h, w), dtype=np.uint8)
        u = rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8)
        v = rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8)
        sws = fo.SwsRgb24(w, h)
        assert np.array_equal(sws(y, u, v), fo.yuv420_to_rgb24(y, u, v))
        sws.close()


def test_cabac_engine_against_literal_spec_decoder(built):
    """csrc/dev/bits.h (32-bit engine scaled by 2^23, 16-bit refills, fused table) against a literal restatement of
    H.264 9.3.3.2 on random bytes and random operation sequences: every bin, every context state and the number of
    consumed bits must agree, from even and odd start positions."""
    import ctypes
    import numpy as np
    from hwang_b200 import build
    L = ctypes.CDLL(build.EMU)
    L.hwb_emu_cabac_selftest.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int]
    rng = np.random.default_rng(5)
    for trial in range(40):
        n = 4096
        # mixtures of random bytes and runs of 0x00 / 0xFF (long MPS / LPS runs, offset near the range boundaries)
        data = rng.integers(0, 256, n, dtype=np.uint8)
        if trial % 3 == 1:
            data[rng.integers(0, n, n // 2)] = 0
        if trial % 3 == 2:
            data[rng.integers(0, n, n // 2)] = 255
        ops = rng.integers(0, 128, 20000, dtype=np.uint8)
        kinds = rng.random(20000)
        ops[kinds < 0.15] = 128          # bypass
        ops[kinds > 0.995] = 129         # terminate (ends the run when it returns 1)
        nctx = [1, 3, 17, 128][trial % 4]
        for start in (0, 1, 7, 16):
            data[start] &= 0x7F  # codIOffset < codIRange at initialisation, as in any conforming stream (9.3.1.2)
            bad = L.hwb_emu_cabac_selftest(data.tobytes() + bytes(16), n, start, ops.tobytes(), len(ops), nctx)
            assert bad == -1, 'engine and spec decoder disagree at operation %d (trial %d, start %d)' % (bad, trial, start)
