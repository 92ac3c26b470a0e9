"""Shared helpers for the parity tests: generate a clip, decode it with the oracle and with our decoder."""
import io
import time

import numpy as np

import hwang_b200 as hw
from hwang_b200.testing import streamgen
from oracle import ffmpeg_oracle as fo


def make_clip(**kw):
    mp4 = streamgen.generate(**kw)
    index = hw.index_video(io.BytesIO(mp4))
    offs, sizes = index.sample_offsets(), index.sample_sizes()
    kf = set(index.keyframe_indices())
    samples = [mp4[o:o + s] for o, s in zip(offs, sizes)]
    return mp4, index, samples, [i in kf for i in range(len(samples))]


def oracle_frames(index, samples, keyflags):
    return fo.decode_samples(index.metadata_bytes(), samples, keyflags)


def decode_yuv(index, samples, keyflags, chunk_pictures=None):
    dec = hw.VideoDecoder(0)
    if chunk_pictures:
        dec.set_chunk_pictures(chunk_pictures)
    dec.configure(index.frame_width(), index.frame_height(), index.format(), index.metadata_bytes())
    for s, k in zip(samples, keyflags):
        dec.feed(s, k)
    dec.feed(None)
    dec.flush()
    out = []
    deadline = time.time() + 300
    while len(out) < len(samples):  # the GPU decodes asynchronously: poll until every fed picture has come out
        n = dec.frames_ready()
        if n != 0:  # negative = decoder error: get_frame_yuv raises with the message
            out.append(dec.get_frame_yuv())
        elif time.time() > deadline:
            raise TimeoutError('decoder produced %d of %d frames' % (len(out), len(samples)))
        else:
            time.sleep(0.0005)
    assert dec.frames_ready() == 0
    return out, dec


def flat(yuv):
    y, u, v = yuv
    return np.concatenate([y.ravel(), u.ravel(), v.ravel()])


def assert_yuv_parity(kw, chunk_pictures=None):
    mp4, index, samples, kf = make_clip(**kw)
    ref = oracle_frames(index, samples, kf)
    got, dec = decode_yuv(index, samples, kf, chunk_pictures)
    assert len(got) == len(ref) == kw['frames']
    for i, (g, r) in enumerate(zip(got, ref)):
        assert np.array_equal(g, flat(r)), 'frame %d differs from libavcodec (%s)' % (i, kw)
    return dec


FEATURE_CLIPS = {
    'cbp_cavlc_multislice_ipcm': dict(frames=20, gop=10, width=320, height=240, seed=11, num_ref=2, slices=3, qp_jitter=3,
                                      ipcm_per_100k=2000, intra_in_p_pct=10),
    'cbp_constrained_intra_poc2': dict(frames=12, gop=12, deblock=3, width=176, height=144, seed=12, num_ref=4, slices=2,
                                       constrained_intra=1, chroma_qp_offset=3, poc_type=2, intra_in_p_pct=20),
    'main_cabac_p': dict(frames=20, gop=10, width=320, height=240, profile=1, seed=21, num_ref=3, slices=2, qp_jitter=3,
                         ipcm_per_100k=2000, intra_in_p_pct=10, cabac_init_idc=-1, deblock=3),
    'high_cabac_b_spatial_scaling': dict(frames=24, gop=12, width=320, height=240, profile=2, seed=31, num_ref=3, slices=2,
                                         qp_jitter=3, ipcm_per_100k=1000, intra_in_p_pct=8, cabac_init_idc=-1, deblock=3,
                                         bframes=2, weighted=2, scaling_lists=1, chroma_qp_offset=-2),
    'high_cavlc_b_temporal': dict(frames=24, gop=12, width=320, height=240, profile=2, seed=32, num_ref=4, slices=3,
                                  qp_jitter=2, intra_in_p_pct=5, bframes=3, direct_spatial=0, weighted=2, cabac=0),
    'main_weighted_p': dict(frames=16, gop=16, width=352, height=288, profile=1, seed=33, num_ref=2, bframes=1, weighted=1, qp=34),
    'deblock_off_and_slice_edges': dict(frames=8, gop=8, width=160, height=128, profile=1, seed=35, slices=4, deblock=2),
    'tiny_16x16': dict(frames=5, gop=5, width=16, height=16, profile=1, seed=36, bframes=1),
    'cropped_1080': dict(frames=4, gop=4, width=1920, height=1080, profile=2, seed=34, num_ref=2, bframes=2, qp=30),
}


# Syntax a real encoder emits that the reference's own clips would exercise (x264: B pyramid, reference list
# modification for duplicate references, MMCO on scene cuts / long-term references): the generator writes it, libavcodec
# decodes it, and both the generator's reconstruction and our decoder must equal libavcodec's output.
SYNTAX_CLIPS = {
    'b_pyramid_spatial': dict(frames=26, gop=13, width=176, height=144, profile=2, seed=101, num_ref=3, bframes=3, b_pyramid=1, weighted=2),
    'b_pyramid_temporal_multislice': dict(frames=26, gop=13, width=176, height=144, profile=2, seed=102, num_ref=4, bframes=3, b_pyramid=1,
                                          direct_spatial=0, weighted=2, slices=2),
    'ref_list_modification_p': dict(frames=20, gop=10, width=176, height=144, profile=1, seed=103, num_ref=4, rplm_pct=70),
    'ref_list_modification_b': dict(frames=24, gop=12, width=176, height=144, profile=2, seed=104, num_ref=4, bframes=2, rplm_pct=70, weighted=2),
    'mmco_long_term_p': dict(frames=40, gop=20, width=176, height=144, profile=1, seed=105, num_ref=3, mmco=1),
    'mmco_long_term_b_temporal': dict(frames=40, gop=20, width=176, height=144, profile=2, seed=106, num_ref=4, bframes=2, mmco=1, rplm_pct=50,
                                      direct_spatial=0),
    'mmco_cavlc_baseline': dict(frames=40, gop=20, width=176, height=144, profile=0, seed=107, num_ref=3, mmco=1, rplm_pct=30),
    'poc_type1_p': dict(frames=20, gop=10, width=176, height=144, profile=1, seed=108, num_ref=2, poc_type=1),
    'poc_type1_b_pyramid': dict(frames=26, gop=13, width=176, height=144, profile=2, seed=109, num_ref=3, bframes=3, b_pyramid=1, poc_type=1),
    'poc_type1_mmco5': dict(frames=40, gop=20, width=176, height=144, profile=1, seed=110, num_ref=3, poc_type=1, mmco=1),
    'poc_type2_mmco5': dict(frames=40, gop=20, width=176, height=144, profile=0, seed=111, num_ref=3, poc_type=2, mmco=1),
    # num_ref_idx_active larger than the number of reference pictures that exist (the first pictures of every GOP): libavcodec
    # lets the missing entries stand for the initial list's first entry, macroblocks refer to them
    'short_ref_lists_p': dict(frames=20, gop=10, width=176, height=144, profile=1, seed=113, num_ref=4, pad_refs=1, weighted=1),
    'short_ref_lists_cavlc_rplm': dict(frames=20, gop=10, width=176, height=144, profile=0, seed=114, num_ref=3, pad_refs=1, rplm_pct=60, slices=2),
    'short_ref_lists_b_temporal_pyramid': dict(frames=26, gop=13, width=176, height=144, profile=2, seed=115, num_ref=4, bframes=3, b_pyramid=1,
                                               pad_refs=1, direct_spatial=0, weighted=2),
    'everything_weighted_p_pyramid': dict(frames=34, gop=17, width=176, height=144, profile=2, seed=112, num_ref=4, bframes=3, b_pyramid=1, mmco=1,
                                          rplm_pct=40, weighted=2, slices=2, qp_jitter=2, intra_in_p_pct=6, scaling_lists=1),
}

# The same kind of clips, added after the last GPU run of round 2: the CPU tier decodes them through the emulation; the GPU
# tier's parameter list (SYNTAX_CLIPS) is left as it was verified on hardware.
CPU_SYNTAX_CLIPS = {
    # I slices inside P / B pictures (intra refresh by slice): slice_type 0..2 instead of 5..7, pictures whose slices differ in type
    'mixed_slice_types_p_cabac': dict(frames=20, gop=10, width=176, height=144, profile=1, seed=116, num_ref=3, slices=3, mixed_slices=1),
    'mixed_slice_types_p_cavlc': dict(frames=20, gop=10, width=176, height=144, profile=0, seed=117, num_ref=2, slices=4, mixed_slices=1, deblock=2),
    'mixed_slice_types_b_temporal_pyramid': dict(frames=26, gop=13, width=176, height=144, profile=2, seed=118, num_ref=4, bframes=3, b_pyramid=1,
                                                 slices=3, mixed_slices=1, direct_spatial=0, weighted=2),
    # explicit weighted prediction in B slices (weighted_bipred_idc = 1): weights and offsets per list and reference
    'explicit_weights_in_b': dict(frames=24, gop=12, width=176, height=144, profile=2, seed=120, num_ref=3, bframes=2, slices=2, weighted=3),
    # headers as other encoders write them: parameter-set ids 3 / 7, pic_init_qp_minus26 = -4, num_ref_idx_default_active = 2
    # ... and 6-bit frame_num / 5-bit pic_order_cnt_lsb: both wrap inside the 75-picture GOPs of the first clip
    'header_variant_p': dict(frames=150, gop=75, width=96, height=80, profile=1, seed=121, num_ref=3, header_variant=1, qp_jitter=2, rplm_pct=30, mmco=1),
    'header_variant_b_cavlc_mmco': dict(frames=72, gop=36, width=176, height=144, profile=2, seed=122, num_ref=4, bframes=2, header_variant=1, cabac=0,
                                        weighted=3, rplm_pct=40, mmco=1),
    'header_variant_b_pyramid_mixed_slices': dict(frames=26, gop=13, width=176, height=144, profile=2, seed=123, num_ref=4, bframes=3, b_pyramid=1,
                                                  header_variant=1, weighted=3, direct_spatial=0, slices=2, mixed_slices=1),
    # long GOPs with a B pyramid, MMCO / long-term references and temporal direct prediction: more than 16 reference pictures per
    # GOP, so frame_num must be wider than 4 bits for libavcodec to be a usable oracle here (it maps co-located references by
    # frame_num: DESIGN.md 6)
    'mmco_long_term_b_temporal_long_gop': dict(frames=80, gop=40, width=80, height=64, profile=2, seed=124, num_ref=3, bframes=2, b_pyramid=1, mmco=1,
                                               direct_spatial=0, header_variant=1, slices=2, qp_jitter=4),
    # direct_8x8_inference_flag = 0: direct prediction per 4x4 block (spatial and temporal), no 8x8 transform in macroblocks with direct parts
    'direct_4x4_spatial_high': dict(frames=24, gop=12, width=96, height=80, profile=2, seed=125, num_ref=3, bframes=2, direct_4x4=1, weighted=2, slices=2, qp_jitter=2),
    'direct_4x4_temporal_pyramid_cavlc': dict(frames=26, gop=13, width=96, height=80, profile=2, seed=126, num_ref=4, bframes=3, b_pyramid=1, direct_4x4=1,
                                              direct_spatial=0, cabac=0, qp_jitter=2),
    'mixed_slice_types_b_spatial_cavlc_constrained': dict(frames=24, gop=12, width=176, height=144, profile=2, seed=119, num_ref=3, bframes=2,
                                                          slices=2, mixed_slices=1, cabac=0, constrained_intra=1, weighted=2),
}


def decode_corrupted_then_clean(seed_list=(1, 2, 3)):
    """Flip bytes inside slice payloads: the decoder must either report an error or return frames, never hang or
    fault, and must decode a clean clip bit-exactly afterwards (same process, same device context)."""
    import random
    kw = dict(width=320, height=240, frames=12, gop=6, profile=1, seed=71, num_ref=2, qp=28)
    mp4, index, samples, kf = make_clip(**kw)
    outcomes = []
    for seed in seed_list:
        rng = random.Random(seed)
        bad = list(samples)
        for victim in (1, 4, 7):  # two P pictures and one picture of the second GOP
            b = bytearray(bad[victim])
            for _ in range(8):
                pos = rng.randrange(16, len(b))  # keep the NAL length field and the slice header mostly intact
                b[pos] ^= 1 << rng.randrange(8)
            bad[victim] = bytes(b)
        dec = hw.VideoDecoder(0)
        dec.configure(kw['width'], kw['height'], index.format(), index.metadata_bytes())
        try:
            for s, k in zip(bad, kf):
                dec.feed(s, k)
            dec.feed(None)
            dec.flush()
            got = 0
            deadline = time.time() + 120
            while got < len(bad) and time.time() < deadline:
                n = dec.frames_ready()
                if n != 0:
                    dec.get_frame_yuv()
                    got += 1
                else:
                    time.sleep(0.0005)
            outcomes.append('frames' if got == len(bad) else 'timeout')
        except RuntimeError as e:
            outcomes.append('error: ' + str(e)[:60])
    assert 'timeout' not in outcomes, outcomes
    assert_yuv_parity(kw)  # the decoder (and the device context) still work
    return outcomes


# ------------------------------------------------------------------------------------------ full-size workloads
# The oracle side of the BASELINE-size parity tests: libavcodec (threads=1 per worker, as the reference runs it) +
# sws_scale in a pool of worker processes, one keyframe-delimited GOP per job, returning one Adler-32 per wanted frame
# (the RGB24 frames themselves would be tens of GB).  The CUDA side computes the same checksum over the bytes it returns.
_POOL_G = {}


def _oracle_gop_job(job):
    import zlib
    a, b, want = job  # absolute sample range of one closed GOP, wanted absolute frame numbers inside it
    nls, sps, pps = fo.parse_avcc(_POOL_G['avcc'])
    w, h = _POOL_G['w'], _POOL_G['h']
    samples = _POOL_G['samples']
    sws = fo.SwsRgb24(w, h)
    dst = np.empty(w * h * 3, np.uint8)
    out = {}
    k = [a]
    wanted = set(want)
    last = max(want)

    def on_frame(addr):
        if k[0] in wanted:
            sws.scale_avframe(addr, dst.ctypes.data)
            out[k[0]] = zlib.adler32(dst)
        k[0] += 1
    dec = fo.FFmpegH264(threads=1)
    for i in range(a, b):
        dec.send_raw(fo.avcc_to_annexb(samples[i], nls, sps, pps, i == a), on_frame)
        if k[0] > last:
            break
    else:
        dec.send_raw(None, on_frame)
    dec.close()
    sws.close()
    return out


def oracle_rgb_checksums(mp4, index, rows, procs=None):
    """{frame number: adler32 of its RGB24 bytes} for the wanted rows, decoded by the reference's ffmpeg path."""
    import multiprocessing as mp
    import os
    offs, sizes = index.sample_offsets(), index.sample_sizes()
    kfs = list(index.keyframe_indices()) + [index.frames()]
    _POOL_G.update(avcc=index.metadata_bytes(), w=index.frame_width(), h=index.frame_height(),
                   samples=[mp4[o:o + s] for o, s in zip(offs, sizes)])
    rows = sorted(rows)
    jobs, j = [], 0
    for a, b in zip(kfs[:-1], kfs[1:]):
        want = []
        while j < len(rows) and rows[j] < b:
            want.append(rows[j])
            j += 1
        if want:
            jobs.append((a, b, want))
    procs = procs or os.cpu_count() or 1
    out = {}
    with mp.get_context('fork').Pool(min(procs, max(1, len(jobs)))) as pool:
        for part in pool.imap_unordered(_oracle_gop_job, jobs):
            out.update(part)
    return out


def automaton_rgb_checksums(index, intervals, total, batch=32, device=0):
    """Adler-32 of every frame DecoderAutomata.get_frames returns for `intervals` (EncodedData list), streamed through one
    page-locked buffer of `batch` frames: [checksum] in request order."""
    import zlib
    from hwang_b200 import _lib
    L = _lib.lib()
    fs = index.frame_width() * index.frame_height() * 3
    auto = hw.DecoderAutomata(hw.DeviceHandle(hw.DeviceType.GPU, device), 1, hw.VideoDecoderType.B200)
    auto.initialize(intervals, index.metadata_bytes())
    pinned = hw.api.PinnedBuffer(fs * batch)
    sums, done = [], 0
    while done < total:
        k = min(batch, total - done)
        if L.hwb_automata_get_frames(auto._h, pinned.ptr, k) != 0:
            raise RuntimeError(L.hwb_automata_last_error(auto._h).decode())
        view = pinned.array[:fs * k].reshape(k, fs)
        sums += [zlib.adler32(view[i]) for i in range(k)]
        done += k
    return sums
