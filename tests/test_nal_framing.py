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


def _scaling_list(b, values, n):
    """scaling_list() syntax (7.3.2.1.1.1): `values` in zig-zag order; a value of 0 at position j > 0 ends the list (the rest
    repeats the last value), a value of 0 at position 0 asks for the default matrix."""
    last = 8
    for j in range(n):
        v = values[j] if j < len(values) else 0
        delta = (v - last + 128) % 256 - 128
        b.se(delta)
        if v == 0:
            return
        last = v


@pytest.mark.parametrize('pps_lists', [False, True])
def test_sequence_level_scaling_matrices_with_fallback_rules(emu, pps_lists):
    """seq_scaling_matrix_present_flag = 1 with every way a list can be given (7.4.2.1.1): explicit, explicit but cut
    short, "use the default matrix", absent -> fall-back rule set A (the previous list, or the default for the first of its
    kind); the PPS carries no matrices of its own, so the sequence-level ones are what dequantisation uses (4x4 and 8x8,
    intra and inter).  The encoder quantised with flat matrices, so the pictures drift -- both decoders must drift alike.
    With pps_lists the picture parameter set carries matrices too (fall-back rule set B).  The matrices stay below twice
    the flat value: with entries of 60 one frame differed from libavcodec -- levels quantised for a flat matrix then
    dequantise beyond the 16 bits a conformant stream guarantees, where libavcodec (16-bit coefficients) and this decoder
    (32-bit) presumably part ways; not pursued."""
    kw = dict(frames=12, gop=6, width=96, height=80, profile=2, seed=64, num_ref=2, bframes=1, intra_in_p_pct=15)
    mp4, index, samples, kf = util.make_clip(**kw)
    avcc = index.metadata_bytes()
    nls, sps_list, pps_list = fo.parse_avcc(avcc)
    b = _Bits()
    b.u(100, 8); b.u(0, 8); b.u(30, 8)
    b.ue(0)                                         # sps id
    b.ue(1); b.ue(0); b.ue(0); b.u(0, 1)            # chroma_format_idc 1, 8 bit, no transform bypass
    b.u(1, 1)                                       # seq_scaling_matrix_present_flag
    b.u(1, 1); _scaling_list(b, [6 + 2 * j for j in range(16)], 16)          # 0: Intra Y, explicit
    b.u(0, 1)                                                                 # 1: Intra Cb <- list 0
    b.u(1, 1); _scaling_list(b, [0], 16)                                      # 2: Intra Cr: default matrix
    b.u(1, 1); _scaling_list(b, [40 - 2 * j for j in range(16)], 16)         # 3: Inter Y, explicit
    b.u(0, 1)                                                                 # 4: Inter Cb <- list 3
    b.u(1, 1); _scaling_list(b, [12, 14, 16, 18, 20], 16)                     # 5: Inter Cr: cut short after five values
    b.u(0, 1)                                                                 # 6: 8x8 Intra: absent -> default
    b.u(1, 1); _scaling_list(b, [9 + (j * 7) % 40 for j in range(64)], 64)    # 7: 8x8 Inter, explicit
    b.ue(0); b.ue(0); b.ue(4)                       # log2_max_frame_num_minus4, poc type 0, 8-bit lsb
    b.ue(2); b.u(0, 1)
    b.ue(5); b.ue(4); b.u(1, 1); b.u(1, 1)
    b.u(0, 1)                                       # no cropping (96x80)
    b.u(1, 1)                                       # VUI: bitstream restriction only
    b.u(0, 1); b.u(0, 1); b.u(0, 1); b.u(0, 1); b.u(0, 1); b.u(0, 1); b.u(0, 1); b.u(0, 1)
    b.u(1, 1); b.u(1, 1); b.ue(0); b.ue(0); b.ue(16); b.ue(16); b.ue(1); b.ue(2)
    sps = bytes([0x67]) + b.rbsp()
    if pps_lists:
        # ... and a picture parameter set with matrices of its own: absent lists fall back to the SEQUENCE-level list
        # (rule set B) for the first of each kind and to the previous list otherwise
        q = _Bits()
        q.ue(0); q.ue(0); q.u(1, 1); q.u(0, 1); q.ue(0)    # ids, CABAC, no field order, one slice group
        q.ue(0); q.ue(0); q.u(0, 1); q.u(0, 2)            # default reference counts, no weighted prediction
        q.se(0); q.se(0); q.se(0)                         # init qp / qs, chroma_qp_index_offset
        q.u(1, 1); q.u(0, 1); q.u(0, 1)                   # deblocking control present, no constrained intra, no redundant pictures
        q.u(1, 1); q.u(1, 1)                              # transform_8x8_mode, pic_scaling_matrix_present
        q.u(0, 1)                                                                 # 0: <- sequence-level list 0 (rule B)
        q.u(1, 1); _scaling_list(q, [26 - j for j in range(16)], 16)              # 1: explicit
        q.u(0, 1)                                                                 # 2: <- list 1
        q.u(0, 1)                                                                 # 3: <- sequence-level list 3 (rule B)
        q.u(1, 1); _scaling_list(q, [0], 16)                                      # 4: default matrix
        q.u(0, 1)                                                                 # 5: <- list 4
        q.u(0, 1)                                                                 # 6: <- sequence-level 8x8 intra
        q.u(1, 1); _scaling_list(q, [12 + (j * 3) % 17 for j in range(64)], 64)   # 7: explicit
        q.se(-1)                                          # second_chroma_qp_index_offset (as tools/h264gen writes it)
        pps_list = [bytes([0x68]) + q.rbsp()]
    new_avcc = bytes(avcc[:5]) + bytes([0xE1]) + struct.pack('>H', len(sps)) + sps + bytes([len(pps_list)]) + \
        b''.join(struct.pack('>H', len(p)) + p for p in pps_list)
    _decode_both(index, new_avcc, samples, kf)
