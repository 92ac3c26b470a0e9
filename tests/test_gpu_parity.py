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


def test_config3_shape_high_bframes_sparse_every_17th(gpu):
    """Config 3 shape at reduced length: 1080p High (8x8 transform, B pictures, implicit weighted bi-prediction),
    rows 0,17,34,... through hwang.Decoder.retrieve; RGB bit-exact against the oracle."""
    kw = dict(width=1920, height=1080, frames=72, gop=24, profile=2, bframes=2, num_ref=3, weighted=2, seed=3, qp=30)
    mp4, index, samples, kf = util.make_clip(**kw)
    ref = util.oracle_frames(index, samples, kf)
    rows = list(range(0, 72, 17))
    frames = hw.Decoder(io.BytesIO(mp4), video_index=index).retrieve(rows)
    assert len(frames) == len(rows)
    for r, f in zip(rows, frames):
        assert np.array_equal(np.asarray(f), fo.yuv420_to_rgb24(*ref[r])), 'row %d' % r


def test_config4_shape_4k_long_gop_random_rows(gpu):
    """Config 4 shape at reduced length: 3840x2160 High, one long GOP, seeded random sparse rows."""
    kw = dict(width=3840, height=2160, frames=20, gop=20, profile=2, bframes=1, num_ref=2, seed=4, qp=32)
    mp4, index, samples, kf = util.make_clip(**kw)
    ref = util.oracle_frames(index, samples, kf)
    rng = np.random.default_rng(0)
    rows = sorted(set(int(x) for x in rng.integers(0, 20, 5)))
    frames = hw.Decoder(io.BytesIO(mp4), video_index=index).retrieve(rows)
    for r, f in zip(rows, frames):
        assert np.array_equal(np.asarray(f), fo.yuv420_to_rgb24(*ref[r])), 'row %d' % r


def test_config5_shape_mixed_resolution_batch(gpu):
    """Config 5 shape at reduced size: clips of several resolutions / profiles decoded by independent automata
    (what each GPU of a sharded run does), every frame checked."""
    from hwang_b200 import shard
    specs = [dict(width=640, height=480, profile=0), dict(width=1280, height=720, profile=1, bframes=1),
             dict(width=1920, height=1080, profile=2, bframes=2), dict(width=320, height=240, profile=1, slices=2)]
    for i, sp in enumerate(specs):
        kw = dict(frames=24, gop=8, seed=50 + i, qp=30, num_ref=2)
        kw.update(sp)
        mp4, index, samples, kf = util.make_clip(**kw)
        ref = util.oracle_frames(index, samples, kf)
        items = shard.gop_work_items(index, i)
        parts = shard.partition(items, 2)
        dec = hw.Decoder(io.BytesIO(mp4), video_index=index)
        seen = 0
        for part in parts:
            for (_, a, b, cost, want) in shard.merge_adjacent(part):
                for r, f in zip(want, dec.retrieve(want)):
                    assert np.array_equal(np.asarray(f), fo.yuv420_to_rgb24(*ref[r])), (i, r)
                    seen += 1
        assert seen == 24


def test_get_frames_in_batches_of_8_and_reconfigure(gpu):
    """The reference GPU test's request shape (decoder_automata_test.cpp:442-445): get_frames in batches of 8."""
    kw = dict(width=352, height=288, frames=64, gop=16, profile=1, bframes=1, seed=61, qp=30)
    mp4, index, samples, kf = util.make_clip(**kw)
    ref = util.oracle_frames(index, samples, kf)
    offs, sizes = index.sample_offsets(), index.sample_sizes()
    ed = hw.EncodedData()
    ed.width, ed.height, ed.format = 352, 288, index.format()
    ed.start_keyframe, ed.end_keyframe = 0, 64
    ed.sample_offsets = [o - offs[0] for o in offs]
    ed.sample_sizes = sizes
    ed.keyframes = index.keyframe_indices()
    ed.valid_frames = list(range(64))
    ed.encoded_video = mp4[offs[0]:offs[-1] + sizes[-1]]
    auto = hw.DecoderAutomata(hw.DeviceHandle(hw.DeviceType.GPU, 0), 1, hw.VideoDecoderType.B200)
    for rep in range(3):
        auto.initialize([ed], index.metadata_bytes())
        got = []
        for _ in range(8):
            got += auto.get_frames(index, 8)
        for r, f in enumerate(got):
            assert np.array_equal(np.asarray(f), fo.yuv420_to_rgb24(*ref[r])), (rep, r)
