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


@pytest.mark.parametrize('name', sorted(util.SYNTAX_CLIPS))
def test_real_encoder_syntax_bit_exact(gpu, name):
    """B pyramid, reference list modification, MMCO / long-term references, POC types 1 and 2 (hwb_testutil.SYNTAX_CLIPS)."""
    util.assert_yuv_parity(util.SYNTAX_CLIPS[name])


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


def test_config2_full_size_properties(gpu):
    """BASELINE configs[1] at full size: the 3000-frame 1080p Main/CABAC GOP-30 clip of bench.py, dense, through
    DecoderAutomata.get_frames.  The oracle cannot decode 3000 frames in seconds, so full size is covered by
    (a) bit-exact RGB against libavcodec + the swscale arithmetic on three whole GOPs (first, middle, last),
    (b) a checksum of per-frame checksums that must not depend on how the clip is cut into chunks (one chunk of
        3000 pictures against chunks of 10 GOPs), which also makes two independent decodes reproduce each other,
    (c) seek == sequential at full size: a sparse request for the last frame of a GOP in the middle of the clip
        returns the bytes the dense pass returned."""
    import zlib
    import bench
    from hwang_b200 import _lib as lib_mod
    mp4 = bench.get_clip(3000)
    index = hw.index_video(io.BytesIO(mp4))
    offs, sizes = index.sample_offsets(), index.sample_sizes()
    kfs = sorted(index.keyframe_indices())
    n = len(offs)
    assert n == 3000 and len(kfs) == 100
    W, H = bench.W, bench.H
    fs = W * H * 3
    L = lib_mod.lib()

    def dense_checksums(chunk_pictures):
        import os
        os.environ['HWB_CHUNK_PICTURES'] = str(chunk_pictures)
        try:
            auto = hw.DecoderAutomata(hw.DeviceHandle(hw.DeviceType.GPU, 0), 1, hw.VideoDecoderType.B200)
        finally:
            del os.environ['HWB_CHUNK_PICTURES']
        ed = hw.EncodedData()
        ed.width, ed.height, ed.format = W, H, index.format()
        ed.start_keyframe, ed.end_keyframe = 0, n
        ed.sample_offsets = [o - offs[0] for o in offs]
        ed.sample_sizes = sizes
        ed.keyframes = kfs
        ed.valid_frames = list(range(n))
        ed.encoded_video = mp4[offs[0]:offs[-1] + sizes[-1]]
        auto.initialize([ed], index.metadata_bytes())
        batch = 50
        pinned = hw.api.PinnedBuffer(fs * batch)
        sums, keep = [], {}
        done = 0
        while done < n:
            k = min(batch, n - done)
            assert L.hwb_automata_get_frames(auto._h, pinned.ptr, k) == 0, L.hwb_automata_last_error(auto._h).decode()
            view = pinned.array[:fs * k].reshape(k, fs)
            for i in range(k):
                sums.append(zlib.adler32(view[i]))
                if (done + i) // 30 in (0, 50, 99):
                    keep[done + i] = view[i].copy()
            done += k
        return sums, keep

    sums_one, keep = dense_checksums(1 << 30)
    sums_cut, _ = dense_checksums(300)
    assert len(sums_one) == n
    assert sums_one == sums_cut, 'result depends on the chunking'
    assert zlib.adler32(np.asarray(sums_one, np.uint32).tobytes()) == zlib.adler32(np.asarray(sums_cut, np.uint32).tobytes())
    # (a) three whole GOPs against the oracle
    for g in (0, 50, 99):
        samples = [mp4[offs[i]:offs[i] + sizes[i]] for i in range(g * 30, g * 30 + 30)]
        ref = fo.decode_samples(index.metadata_bytes(), samples, [i == 0 for i in range(30)])
        for i in range(30):
            exp = fo.yuv420_to_rgb24(*ref[i]).reshape(-1)
            assert np.array_equal(keep[g * 30 + i], exp), 'frame %d differs from the oracle' % (g * 30 + i)
    # (c) seek == sequential
    single = hw.Decoder(io.BytesIO(mp4), video_index=index).retrieve([50 * 30 + 29])
    assert np.array_equal(np.asarray(single[0]).reshape(-1), keep[50 * 30 + 29])


def test_corrupted_payload_is_survivable(gpu):
    """Bit flips inside slice payloads: an error or garbage frames, never a hang or a device fault; a clean clip
    decodes bit-exactly afterwards in the same process."""
    outcomes = util.decode_corrupted_then_clean()
    assert len(outcomes) == 3
