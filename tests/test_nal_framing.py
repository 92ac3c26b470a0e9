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


class _Bits:
    def __init__(self): self.bits = []
    def u(self, v, n): self.bits += [(v >> (n - 1 - i)) & 1 for i in range(n)]
    def ue(self, v):
        v += 1; n = v.bit_length()
        self.u(0, n - 1); self.u(v, n)
    def se(self, v): self.ue(2 * v - 1 if v > 0 else -2 * v)
    def rbsp(self):
        b = self.bits + [1]
        b += [0] * (-len(b) % 8)
        raw = bytes(int(''.join(map(str, b[i:i + 8])), 2) for i in range(0, len(b), 8))
        out, zeros = bytearray(), 0
        for x in raw:  # emulation prevention
            if zeros >= 2 and x <= 3:
                out.append(3); zeros = 0
            out.append(x)
            zeros = zeros + 1 if x == 0 else 0
        return bytes(out)


def test_sps_with_full_vui_hrd_and_timing(emu):
    """The sequence parameter set as x264 --nal-hrd / hardware encoders write it: sample aspect ratio, overscan, video
    signal type with colour description, chroma location, timing info, NAL and VCL HRD parameters, bitstream
    restriction -- everything ahead of the fields a decoder needs must be skipped bit-exactly."""
    kw = dict(frames=12, gop=6, width=96, height=72, profile=1, seed=62, num_ref=2, bframes=2)  # 72 = 80 - 8: cropping too
    mp4, index, samples, kf = util.make_clip(**kw)
    avcc = index.metadata_bytes()
    nls, sps_list, pps_list = fo.parse_avcc(avcc)
    mb_w, mb_h = 6, 5
    b = _Bits()
    b.u(77, 8); b.u(0x40, 8); b.u(30, 8)            # Main, constraint_set1, level 3.0 (as tools/h264gen writes them)
    b.ue(0)                                         # sps id
    b.ue(0)                                         # log2_max_frame_num_minus4
    b.ue(0); b.ue(4)                                # poc type 0, 8-bit lsb
    b.ue(2); b.u(0, 1)                              # max_num_ref_frames, gaps
    b.ue(mb_w - 1); b.ue(mb_h - 1); b.u(1, 1); b.u(1, 1)
    b.u(1, 1); b.ue(0); b.ue(0); b.ue(0); b.ue(4)   # crop bottom 8 rows
    b.u(1, 1)                                       # vui_parameters_present
    b.u(1, 1); b.u(255, 8); b.u(4, 16); b.u(3, 16)  # aspect_ratio: Extended_SAR 4:3
    b.u(1, 1); b.u(1, 1)                            # overscan info
    b.u(1, 1); b.u(5, 3); b.u(0, 1); b.u(1, 1); b.u(1, 8); b.u(1, 8); b.u(1, 8)  # video signal type + colour description
    b.u(1, 1); b.ue(0); b.ue(0)                     # chroma loc
    b.u(1, 1); b.u(1001, 32); b.u(60000, 32); b.u(1, 1)  # timing info (contains 00 00 01/03 patterns: emulation prevention)
    for _ in range(2):                              # nal_hrd, vcl_hrd
        b.u(1, 1); b.ue(1); b.u(4, 4); b.u(6, 4)
        for _ in range(2): b.ue(7811); b.ue(1561); b.u(0, 1)
        b.u(23, 5); b.u(23, 5); b.u(23, 5); b.u(24, 5)
    b.u(0, 1)                                       # low_delay_hrd
    b.u(0, 1)                                       # pic_struct_present
    b.u(1, 1); b.u(1, 1); b.ue(0); b.ue(0); b.ue(11); b.ue(11); b.ue(1); b.ue(3)  # bitstream restriction: 1 reorder frame
    sps = bytes([0x67]) + b.rbsp()
    new_avcc = bytes(avcc[:5]) + bytes([0xE1]) + struct.pack('>H', len(sps)) + sps + bytes([len(pps_list)]) + \
        b''.join(struct.pack('>H', len(p)) + p for p in pps_list)
    _decode_both(index, new_avcc, samples, kf)


def test_frame_cropping_with_a_top_offset(emu):
    """frame_crop_top_offset other than 0 (the generator only crops at the right and at the bottom): the same coded
    pictures shown through a window that starts at luma row 4 -- planar output and RGB24 against libavcodec + the
    swscale arithmetic.  (A left offset is not compared: libavcodec itself reduces frame_crop_left_offset to a multiple
    of 32 samples "to preserve alignment", h264_ps.c, so the reference would show a 92-sample-wide picture there.)"""
    kw = dict(frames=8, gop=4, width=88, height=72, profile=1, seed=63, num_ref=2, bframes=1)  # coded 96x80
    mp4, index, samples, kf = util.make_clip(**kw)
    avcc = index.metadata_bytes()
    nls, sps_list, pps_list = fo.parse_avcc(avcc)
    b = _Bits()
    b.u(77, 8); b.u(0x40, 8); b.u(30, 8)
    b.ue(0); b.ue(0); b.ue(0); b.ue(4)
    b.ue(2); b.u(0, 1)
    b.ue(5); b.ue(4); b.u(1, 1); b.u(1, 1)
    b.u(1, 1); b.ue(0); b.ue(4); b.ue(2); b.ue(2)   # crop 8 luma samples on the right, 4 rows at the top and at the bottom: 96x80 -> 88x72
    b.u(0, 1)                                       # no VUI at all (output order must not depend on it)
    sps = bytes([0x67]) + b.rbsp()
    new_avcc = bytes(avcc[:5]) + bytes([0xE1]) + struct.pack('>H', len(sps)) + sps + bytes([len(pps_list)]) + \
        b''.join(struct.pack('>H', len(p)) + p for p in pps_list)
    _decode_both(index, new_avcc, samples, kf)
    ref = fo.decode_samples(new_avcc, samples, kf)
    dec = hw.VideoDecoder(0)
    dec.configure(88, 72, index.format(), new_avcc)
    for s, k in zip(samples, kf):
        dec.feed(s, k)
    dec.feed(None); dec.flush()
    for i in range(len(samples)):
        while dec.frames_ready() == 0:
            pass
        rgb = np.asarray(dec.get_frame()).reshape(72, 88, 3)
        assert np.array_equal(rgb, fo.yuv420_to_rgb24(*ref[i])), 'RGB24 frame %d' % i
