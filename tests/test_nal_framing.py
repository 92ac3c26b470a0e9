"""Sample framing as real muxers write it (CPU tier: host parser + emulated device against libavcodec): NAL length
fields of 1, 2 or 3 bytes instead of 4 (avcC lengthSizeMinusOne), access unit delimiters, SEI and filler NAL units and
zero-length NAL units inside a sample, parameter sets repeated in band at key frames.  The reference hands all of that
to h264_mp4toannexb + libavcodec (software_video_decoder.cpp:173-202); the oracle does the same."""
import io
import struct

import numpy as np
import pytest

import hwang_b200 as hw
import hwb_testutil as util
from oracle import ffmpeg_oracle as fo

KW = dict(frames=10, gop=5, width=96, height=80, profile=1, seed=61, num_ref=2, slices=2)


def _nals(sample, nls=4):
    out, off = [], 0
    while off + nls <= len(sample):
        n = int.from_bytes(sample[off:off + nls], 'big'); off += nls
        out.append(sample[off:off + n]); off += n
    return out


def _frame(nals, nls):
    return b''.join(len(n).to_bytes(nls, 'big') + n for n in nals)


def _decode_both(index, avcc, samples, kf):
    ref = fo.decode_samples(avcc, samples, kf)
    dec = hw.VideoDecoder(0)
    dec.configure(index.frame_width(), index.frame_height(), index.format(), avcc)
    for s, k in zip(samples, kf):
        dec.feed(s, k)
    dec.feed(None); dec.flush()
    got = []
    while len(got) < len(samples):
        assert dec.frames_ready() != 0 or len(got) < len(samples)
        if dec.frames_ready() != 0:
            got.append(dec.get_frame_yuv())
    assert len(ref) == len(got) == len(samples)
    for i, (g, r) in enumerate(zip(got, ref)):
        assert np.array_equal(g, util.flat(r)), 'frame %d differs from libavcodec' % i


@pytest.mark.parametrize('nls', [1, 2, 3])
def test_short_nal_length_fields(emu, nls):
    kw = dict(KW)
    if nls == 1:
        kw.update(width=48, height=32, slices=4, qp=34)  # every NAL unit must fit 255 bytes
    mp4, index, samples, kf = util.make_clip(**kw)
    avcc = bytearray(index.metadata_bytes())
    assert (avcc[4] & 3) == 3
    avcc[4] = (avcc[4] & 0xFC) | (nls - 1)
    re = []
    for s in samples:
        nals = _nals(s)
        if any(len(n) >= 1 << (8 * nls) for n in nals):
            pytest.skip('a NAL unit of this clip does not fit a %d-byte length field' % nls)
        re.append(_frame(nals, nls))
    _decode_both(index, bytes(avcc), re, kf)


def test_aud_sei_filler_empty_and_inband_parameter_sets(emu):
    mp4, index, samples, kf = util.make_clip(**KW)
    avcc = index.metadata_bytes()
    _, sps, pps = fo.parse_avcc(avcc)
    aud = bytes([0x09, 0xF0])                                   # access unit delimiter, primary_pic_type 7
    sei = bytes([0x06, 0x05, 0x04, 1, 2, 3, 4, 0x80])           # user_data_unregistered-shaped payload (ignored by both decoders)
    filler = bytes([0x0C, 0xFF, 0xFF, 0xFF, 0x80])              # filler data
    re = []
    for i, (s, k) in enumerate(zip(samples, kf)):
        nals = _nals(s)
        pre = [aud]
        if k:
            pre += list(sps) + list(pps)                         # parameter sets repeated in band at key frames
        pre += [sei, b'']                                        # a zero-length NAL unit is skipped
        post = [filler] if i % 2 else []
        re.append(_frame(pre + nals + post, 4))
    _decode_both(index, avcc, re, kf)
