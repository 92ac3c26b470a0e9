"""GPU parity tests: the CUDA path, called through the C ABI, against the libavcodec oracle."""
import io

import numpy as np
import pytest

import hwang_b200 as hw
from oracle import ffmpeg_oracle as fo
import hwb_testutil as util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', sorted(util.FEATURE_CLIPS))
def test_yuv_bit_exact(gpu, name):
    util.assert_yuv_parity(util.FEATURE_CLIPS[name])


def test_small_chunks_match(gpu):
    # chunk boundaries at every GOP: results must not depend on batching
    util.assert_yuv_parity(dict(frames=30, gop=5, width=176, height=144, profile=1, seed=41, bframes=1), chunk_pictures=1)


def test_retrieve_rgb_sparse_config1(gpu):
    """Config 1 shape: 640x480 Constrained Baseline CAVLC GOP 30, the reference test's wanted set
    (decoder_automata_test.cpp:233-245) plus the frame-250 gather (:287), through hwang.Decoder.retrieve."""
    kw = dict(width=640, height=480, frames=300, gop=30, profile=0, seed=1, qp=28)
    mp4, index, samples, kf = util.make_clip(**kw)
    ref = util.oracle_frames(index, samples, kf)
    rows = list(range(0, 10)) + list(range(30, 55)) + list(range(100, 120)) + list(range(121, 170)) + [250]
    dec = hw.Decoder(io.BytesIO(mp4), video_index=index)
    frames = dec.retrieve(rows)
    assert len(frames) == len(rows)
    sws = fo.SwsRgb24(640, 480)
    for r, f in zip(rows, frames):
        assert np.array_equal(np.asarray(f), fo.yuv420_to_rgb24(*ref[r])), 'row %d' % r
    # the formula restatement itself is pinned to the real swscale on a few frames
    for r in (0, 121, 250):
        assert np.array_equal(sws(*ref[r]), fo.yuv420_to_rgb24(*ref[r]))
    # seek == sequential (the reference's only pixel assertion, decoder_automata_test.cpp:338-340)
    single = hw.Decoder(io.BytesIO(mp4), video_index=index).retrieve([250])
    assert np.array_equal(np.asarray(single[0]), np.asarray(frames[-1]))
