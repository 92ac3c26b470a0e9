"""CPU tier: the C ABI library loads and exports every symbol include/hwang_b200.h declares; automaton logic."""
import io
import os
import re

import numpy as np
import pytest

import hwang_b200 as hw
from hwang_b200 import _lib
from oracle import ffmpeg_oracle as fo
import hwb_testutil as util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, 'include', 'hwang_b200.h')) as f:
        src = re.sub(r'/\*.*?\*/', '', f.read(), flags=re.S)
    return sorted(set(re.findall(r'\b(hwb_[a-z0-9_]+)\s*\(', src)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.SIGNATURES)


def test_product_library_exports_every_declared_symbol():
    """No compute call: on the CPU box this only checks that the CUDA-linked library loads and exports the ABI."""
    import ctypes
    from hwang_b200 import build
    build.build_product()
    lib = ctypes.CDLL(_lib.PRODUCT_LIB)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    lib.hwb_version.restype = ctypes.c_char_p
    assert b'sm_100a' in lib.hwb_version()


def test_product_has_no_cpu_fallback():
    """Without a CUDA device the product library must refuse to build a decoder (never decode on the CPU)."""
    import ctypes
    lib = ctypes.CDLL(_lib.PRODUCT_LIB)
    if lib.hwb_device_count() > 0:
        pytest.skip('a GPU is present')
    h = ctypes.c_void_p()
    assert lib.hwb_decoder_create(1, 0, 1, 3, ctypes.byref(h)) != 0 and not h
    lib.hwb_automata_create.restype = ctypes.c_void_p
    assert not lib.hwb_automata_create(1, 0, 1, 3)


def test_factory_rejects_other_backends(emu):
    L = _lib.lib()
    assert L.hwb_has_decoder_type(hw.VideoDecoderType.B200) == 1
    assert L.hwb_has_decoder_type(hw.VideoDecoderType.SOFTWARE) == 0
    with pytest.raises(RuntimeError):
        hw.DecoderAutomata(hw.DeviceHandle(hw.DeviceType.CPU, 0), 1, hw.VideoDecoderType.SOFTWARE)


def test_unsupported_codec_and_bad_extradata(emu):
    dec = hw.VideoDecoder(0)
    with pytest.raises(RuntimeError, match='Unsupported video codec'):
        dec.configure(64, 48, 'hev1', b'\x01' * 16)
    with pytest.raises(RuntimeError):
        dec.configure(64, 48, 'avc1', b'\x00' * 4)


def test_sparse_retrieval_shapes(emu, built):
    """The reference tests' request shapes (decoder_automata_test.cpp:233-245, :287, :442-445) on a small clip:
    ranges across GOPs, a single-frame gather, get_frames in batches, several intervals in one initialize."""
    kw = dict(width=96, height=80, frames=60, gop=6, profile=1, bframes=1, seed=77, qp=30)
    mp4, index, samples, kf = util.make_clip(**kw)
    ref = [fo.yuv420_to_rgb24(*f) for f in util.oracle_frames(index, samples, kf)]
    dec = hw.Decoder(io.BytesIO(mp4), video_index=index)
    for rows in ([0], [59], [25], list(range(0, 10)) + list(range(30, 55)), list(range(0, 60, 17)), list(range(60)), [5, 6, 7, 41]):
        frames = dec.retrieve(rows)
        assert len(frames) == len(rows)
        for r, f in zip(rows, frames):
            assert np.array_equal(np.asarray(f), ref[r]), (rows, r)
    # all intervals handed to one initialize, frames pulled in batches that straddle interval boundaries
    rows = [1, 2, 13, 14, 15, 40, 58]
    ivs = hw.slice_into_video_intervals(index, rows)
    offs, sizes = index.sample_offsets(), index.sample_sizes()
    eds = []
    for (s, e), valid in ivs:
        d = hw.EncodedData()
        d.width, d.height, d.format = kw['width'], kw['height'], index.format()
        d.start_keyframe, d.end_keyframe = s, e
        d.sample_offsets = [o - offs[s] for o in offs[s:e]]
        d.sample_sizes = sizes[s:e]
        d.keyframes = [k for k in index.keyframe_indices() if s <= k <= e]
        d.valid_frames = valid
        d.encoded_video = mp4[offs[s]:offs[e - 1] + sizes[e - 1]]
        eds.append(d)
    assert len(eds) > 1
    auto = hw.DecoderAutomata(hw.DeviceHandle(hw.DeviceType.GPU, 0), 1, hw.VideoDecoderType.B200)
    auto.initialize(eds, index.metadata_bytes())
    got = []
    for n in (3, 1, 3):
        got += auto.get_frames(index, n)
    for r, f in zip(rows, got):
        assert np.array_equal(np.asarray(f), ref[r])
    with pytest.raises(RuntimeError):
        auto.get_frames(index, 1)  # more than the intervals hold
    # re-initialize resets everything
    auto.initialize(eds[:1], index.metadata_bytes())
    f = auto.get_frames(index, 1)
    assert np.array_equal(np.asarray(f[0]), ref[rows[0]])


def test_corrupt_stream_reports_error(emu, built):
    kw = dict(width=64, height=48, frames=6, gop=6, profile=1, seed=9)
    mp4, index, samples, kf = util.make_clip(**kw)
    dec = hw.VideoDecoder(0)
    dec.configure(64, 48, 'avc1', index.metadata_bytes())
    with pytest.raises(RuntimeError):
        dec.feed(samples[1], False)  # interval must start at an IDR
    dec.configure(64, 48, 'avc1', index.metadata_bytes())
    with pytest.raises(RuntimeError):
        dec.feed(samples[0][:20], True)  # NAL length runs past the sample


def test_corrupted_payload_is_survivable_host_emulation(emu, built):
    outcomes = util.decode_corrupted_then_clean()
    assert len(outcomes) == 3


def test_batch_retrieval_across_workers(emu, built):
    """hwang_b200.batch.retrieve_many: several clips, sparse rows, two workers (both on the emulated device): the
    frames equal what a single Decoder returns for the same rows."""
    from hwang_b200 import batch
    clips = [dict(width=64, height=48, frames=12, gop=4, profile=1, seed=81, bframes=1),
             dict(width=96, height=64, frames=9, gop=3, profile=0, seed=82),
             dict(width=64, height=64, frames=10, gop=5, profile=2, seed=83, bframes=2)]
    reqs, want = [], []
    for i, kw in enumerate(clips):
        mp4, index, samples, kf = util.make_clip(**kw)
        rows = [0, 2, 5, kw['frames'] - 1] if i != 1 else list(range(kw['frames']))
        reqs.append((mp4, rows))
        want.append(hw.Decoder(io.BytesIO(mp4), video_index=index).retrieve(sorted(set(rows))))
    got = batch.retrieve_many(reqs, devices=[0, 0])
    assert len(got) == len(clips)
    for g, w in zip(got, want):
        assert len(g) == len(w)
        for a, b in zip(g, w):
            assert np.array_equal(np.asarray(a), np.asarray(b))


def test_feeder_stops_after_the_last_wanted_frame(emu, built):
    """The tail of a GOP after the last wanted frame is not decoded (reference: decoder_automata.cpp:287).  Long GOP
    with B pictures, one early row per request: the returned frame is bit-exact and far fewer pictures than the GOP
    holds were decoded; rows at the very end of the GOP still decode all of it."""
    kw = dict(width=96, height=64, frames=80, gop=40, profile=2, bframes=3, num_ref=2, seed=91, qp=30)
    mp4, index, samples, kf = util.make_clip(**kw)
    ref = [fo.yuv420_to_rgb24(*f) for f in util.oracle_frames(index, samples, kf)]
    # (the reference's interval slicer starts row 41's interval at keyframe 0: the two GOPs are byte-adjacent,
    # video_index.cpp:76-84 -- so 42 + 16 pictures are fed there, still short of the 80 the interval holds)
    for row, max_decoded in ((1, 18), (5, 22), (41, 58), (39, 40), (79, 80)):
        dec = hw.Decoder(io.BytesIO(mp4), video_index=index)
        s0 = dec._decoder.stats()['pictures_decoded']
        f = dec.retrieve([row])
        assert np.array_equal(np.asarray(f[0]), ref[row]), row
        assert dec._decoder.stats()['pictures_decoded'] - s0 <= max_decoded, row
    # several rows spread over both GOPs in one call
    rows = [0, 3, 9, 44, 52]
    frames = hw.Decoder(io.BytesIO(mp4), video_index=index).retrieve(rows)
    for r, f in zip(rows, frames):
        assert np.array_equal(np.asarray(f), ref[r]), r


def test_unrequested_non_reference_pictures_are_skipped(emu, built):
    """Sparse rows over a clip with non-reference B pictures: every returned frame is bit-exact, and the pictures
    that were neither requested nor referenced were not decoded at all (the reference decodes and drops them,
    decoder_automata.cpp:235; SURVEY 8a row A2 allows a backend to skip them)."""
    kw = dict(width=96, height=64, frames=48, gop=24, profile=2, bframes=2, num_ref=2, seed=92, qp=30)
    mp4, index, samples, kf = util.make_clip(**kw)
    ref = [fo.yuv420_to_rgb24(*f) for f in util.oracle_frames(index, samples, kf)]
    # dense: nothing may be skipped
    dec = hw.Decoder(io.BytesIO(mp4), video_index=index)
    frames = dec.retrieve(list(range(48)))
    assert dec._decoder.stats()['pictures_decoded'] == 48
    for r, f in enumerate(frames):
        assert np.array_equal(np.asarray(f), ref[r]), r
    # every 5th frame: B pictures that are not requested are left out
    rows = list(range(0, 48, 5))
    dec = hw.Decoder(io.BytesIO(mp4), video_index=index)
    frames = dec.retrieve(rows)
    for r, f in zip(rows, frames):
        assert np.array_equal(np.asarray(f), ref[r]), r
    decoded = dec._decoder.stats()['pictures_decoded']
    assert decoded < 40, decoded  # 48 fed; two of every three pictures are non-reference B pictures
    # a plain VideoDecoder (no automaton, no hint) decodes everything
    got, d2 = util.decode_yuv(index, samples, kf)
    assert d2.stats()['pictures_decoded'] == 48


class _Bits:
    def __init__(self):
        self.b = []

    def u(self, n, v):
        self.b += [(v >> (n - 1 - i)) & 1 for i in range(n)]

    def ue(self, v):
        v += 1
        n = v.bit_length()
        self.u(n - 1, 0)
        self.u(n, v)

    def se(self, v):
        self.ue(2 * v - 1 if v > 0 else -2 * v)

    def rbsp(self):
        bits = self.b + [1]
        bits += [0] * (-len(bits) % 8)
        raw = bytes(int(''.join(map(str, bits[i:i + 8])), 2) for i in range(0, len(bits), 8))
        out = bytearray()
        zeros = 0
        for x in raw:  # emulation prevention
            if zeros >= 2 and x <= 3:
                out.append(3)
                zeros = 0
            out.append(x)
            zeros = zeros + 1 if x == 0 else 0
        return bytes(out)


def _sps(profile=66, chroma=1, depth=8, frame_mbs_only=1, mbw=4, mbh=3):
    b = _Bits()
    b.u(8, profile); b.u(8, 0xC0 if profile == 66 else 0); b.u(8, 30)
    b.ue(0)
    if profile >= 100:
        b.ue(chroma)
        if chroma == 3:
            b.u(1, 0)
        b.ue(depth - 8); b.ue(depth - 8); b.u(1, 0); b.u(1, 0)  # bit depths, qpprime bypass, no scaling matrix
    b.ue(0); b.ue(2)            # log2_max_frame_num_minus4, pic_order_cnt_type 2
    b.ue(1); b.u(1, 0)          # max_num_ref_frames, gaps
    b.ue(mbw - 1); b.ue(mbh - 1)
    b.u(1, frame_mbs_only)
    if not frame_mbs_only:
        b.u(1, 0)               # mb_adaptive_frame_field_flag
    b.u(1, 1); b.u(1, 0); b.u(1, 0)  # direct_8x8_inference, no cropping, no VUI
    return bytes([0x67]) + b.rbsp()


def _pps(slice_groups=1, cabac=0):
    b = _Bits()
    b.ue(0); b.ue(0); b.u(1, cabac); b.u(1, 0)
    b.ue(slice_groups - 1)
    if slice_groups > 1:
        b.ue(0)                 # slice_group_map_type 0: run lengths
        for _ in range(slice_groups):
            b.ue(0)
    b.ue(0); b.ue(0); b.u(1, 0); b.u(2, 0); b.se(0); b.se(0); b.se(0); b.u(1, 1); b.u(1, 0); b.u(1, 0)
    return bytes([0x68]) + b.rbsp()


def _avcc(sps, pps):
    return bytes([1, sps[1], sps[2], sps[3], 0xFF, 0xE1]) + len(sps).to_bytes(2, 'big') + sps + bytes([1]) + len(pps).to_bytes(2, 'big') + pps


def test_unsupported_stream_features_are_refused_at_configure(emu):
    """The supported subset is enforced with an error Result, never with silent garbage."""
    dec = hw.VideoDecoder(0)
    dec.configure(64, 48, 'avc1', _avcc(_sps(), _pps()))  # the hand-made parameter sets themselves are fine
    for avcc, msg in ((_avcc(_sps(frame_mbs_only=0), _pps()), 'interlaced'),
                      (_avcc(_sps(profile=100, chroma=2), _pps()), '4:2:0 8-bit'),
                      (_avcc(_sps(profile=100, depth=10), _pps()), '4:2:0 8-bit'),
                      (_avcc(_sps(), _pps(slice_groups=2)), 'FMO')):
        with pytest.raises(RuntimeError, match=msg):
            hw.VideoDecoder(0).configure(64, 48, 'avc1', avcc)
    with pytest.raises(RuntimeError, match='does not match the SPS'):
        hw.VideoDecoder(0).configure(128, 48, 'avc1', _avcc(_sps(), _pps()))


def test_many_batches_device_pointers_and_reuse(emu, built):
    """Several GPU batches in flight (one GOP each), frames fetched as borrowed device pointers, then the same decoder
    re-configured and used again: exercises batch retirement / memory recycling (the emulated device's pointers are
    host pointers, so the borrowed frames can be read directly)."""
    import ctypes
    kw = dict(width=64, height=48, frames=20, gop=4, profile=1, seed=91, bframes=1)
    mp4, index, samples, kf = util.make_clip(**kw)
    ref = util.oracle_frames(index, samples, kf)
    dec = hw.VideoDecoder(0)
    dec.set_chunk_pictures(1)
    fs = 64 * 48 * 3
    for rep in range(3):
        dec.configure(64, 48, index.format(), index.metadata_bytes())
        for s, k in zip(samples, kf):
            dec.feed(s, k)
        dec.feed(None)
        dec.flush()
        ptrs = [dec.get_frame_device() for _ in range(len(samples))]
        for i, p in enumerate(ptrs):
            got = np.ctypeslib.as_array((ctypes.c_uint8 * fs).from_address(p)).reshape(48, 64, 3)
            assert np.array_equal(got, fo.yuv420_to_rgb24(*ref[i])), (rep, i)
        dec.wait_until_frames_copied()
        assert dec.stats()['chunks'] == 5 * (rep + 1)


def test_device_frames_and_oracle_checksum_helpers(emu, built):
    """Decoder.retrieve_device on the emulated device (its "device memory" is host memory) and the full-size tests'
    oracle helper on a small clip: both agree with the per-frame oracle."""
    import zlib
    kw = dict(width=96, height=64, frames=24, gop=6, profile=2, bframes=2, b_pyramid=0, num_ref=3, weighted=2, seed=95)
    mp4, index, samples, kf = util.make_clip(**kw)
    ref = util.oracle_frames(index, samples, kf)
    rows = [0, 3, 5, 6, 17, 23]
    dev = hw.Decoder(io.BytesIO(mp4), video_index=index).retrieve_device(rows)
    host = dev.to_host()
    for k, r in enumerate(rows):
        assert np.array_equal(host[k], fo.yuv420_to_rgb24(*ref[r])), r
    sums = util.oracle_rgb_checksums(mp4, index, rows, procs=2)
    assert sums == {r: zlib.adler32(fo.yuv420_to_rgb24(*ref[r]).reshape(-1)) for r in rows}


def test_batch_retrieval_collects_clips_into_common_batches(emu, built):
    """retrieve_many: intervals of several clips of equal geometry go through one decoder with deferred submission, so
    they share GPU batches: 3 clips x 4 intervals decode in fewer batches than intervals."""
    from hwang_b200 import batch
    reqs, refs = [], []
    for i in range(3):
        mp4, index, samples, kf = util.make_clip(width=64, height=48, frames=16, gop=4, profile=1, seed=120 + i, bframes=i % 2)
        refs.append(util.oracle_frames(index, samples, kf))
        reqs.append((mp4, [1, 6, 9, 15]))
    got = batch.retrieve_many(reqs, devices=[0])
    for frames, ref in zip(got, refs):
        for r, f in zip([1, 6, 9, 15], frames):
            assert np.array_equal(np.asarray(f), fo.yuv420_to_rgb24(*ref[r]))


def test_pinned_buffer_pool_reuses_and_bounds(emu):
    """api.PinnedBuffer: a released buffer serves the next request of a similar size, the pool is bounded, and
    release_pinned_pool() hands everything back (the emulation's 'page-locked' allocator is malloc)."""
    from hwang_b200 import api
    api.release_pinned_pool()
    a = api.PinnedBuffer(3 << 20)
    pa, cap = a.ptr, a.capacity
    assert cap >= 3 << 20 and len(a.array) == 3 << 20
    a.array[:16] = 7
    del a
    assert api._pool_bytes == cap
    b = api.PinnedBuffer((3 << 20) - 4096)  # similar size: reused
    assert b.ptr == pa and api._pool_bytes == 0
    c = api.PinnedBuffer(64 << 20)          # much larger: a new allocation
    assert c.ptr != pa
    del b, c
    assert api._pool_bytes == cap + (64 << 20)
    small = api.PinnedBuffer(1 << 10)       # far smaller than anything idle: not served from a 64 MB buffer
    assert small.capacity <= 2 * (1 << 10) + api._POOL_GRAIN
    del small
    old_max = api._POOL_MAX
    try:
        api._POOL_MAX = 16 << 20            # shrink the bound: the next release evicts the oldest buffers
        d = api.PinnedBuffer(5 << 20)
        del d
        assert api._pool_bytes <= 16 << 20
    finally:
        api._POOL_MAX = old_max
    api.release_pinned_pool()
    assert api._pool_bytes == 0 and not api._pool
