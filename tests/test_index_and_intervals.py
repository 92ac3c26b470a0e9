"""CPU tier: MP4IndexCreator / VideoIndex / slice_into_video_intervals against the reference's own outputs
(tests/golden/*.json, produced by oracle/_ref/ref_tool) and the oracle restatements."""
import glob
import io
import json
import os
import struct

import numpy as np
import pytest

import hwang_b200 as hw
from hwang_b200.testing import streamgen
from oracle import intervals as oiv, mp4_simple

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
NAMES = sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(GOLDEN, '*.json')))


def load(name):
    with open(os.path.join(GOLDEN, name + '.mp4'), 'rb') as f:
        mp4 = f.read()
    with open(os.path.join(GOLDEN, name + '.json')) as f:
        return mp4, json.load(f)


def run_indexer(mp4, step=1024):
    """the 1 KiB pull loop of hwang/mp4_index_creator_test.cpp:36-41"""
    ic = hw.MP4IndexCreator(len(mp4))
    off, size = 0, step
    guard = 0
    while not ic.is_done():
        _, off, size = ic.feed(mp4[off:off + size], size)
        guard += 1
        assert guard < 100000
    return ic


@pytest.mark.parametrize('name', NAMES)
def test_index_matches_reference_index_creator(emu, name):
    mp4, g = load(name)
    ic = run_indexer(mp4)
    assert not ic.is_error(), ic.error_message()
    vi = ic.get_video_index()
    ri = g['reference_index']
    assert vi.sample_offsets() == ri['offsets']
    assert vi.sample_sizes() == ri['sizes']
    assert vi.keyframe_indices() == ri['keyframes']
    assert vi.metadata_bytes().hex() == ri['metadata_hex']
    assert (vi.frame_width(), vi.frame_height(), vi.format(), vi.frames()) == (ri['width'], ri['height'], ri['format'], ri['frames'])
    assert (vi.timescale(), vi.duration()) == (ri['timescale'], ri['duration'])
    # and the python oracle walker agrees
    o = mp4_simple.index_mp4(mp4)
    assert (o['offsets'], o['sizes'], o['keyframes']) == (ri['offsets'], ri['sizes'], ri['keyframes'])


@pytest.mark.parametrize('name', NAMES)
def test_slice_into_video_intervals_matches_reference(emu, name):
    mp4, g = load(name)
    vi = run_indexer(mp4).get_video_index()
    for case in g['reference_intervals']:
        got = hw.slice_into_video_intervals(vi, case['rows'])
        assert got == [((d['start'], d['end']), d['rows']) for d in case['intervals']]


def test_intervals_property_random(emu):
    rng = np.random.default_rng(5)
    for trial in range(200):
        n = int(rng.integers(1, 200))
        gop = int(rng.integers(1, 20))
        kfs = list(range(0, n, gop))
        sizes = [int(x) for x in rng.integers(1, 50, n)]
        offs, o = [], 100
        for i in range(n):
            if i in kfs and rng.random() < 0.3:
                o += int(rng.integers(1, 10))  # gap: GOP not byte-adjacent
            offs.append(o)
            o += sizes[i]
        rows = sorted(set(int(x) for x in rng.integers(0, n, int(rng.integers(1, 40)))))
        vi = hw.VideoIndex.create(30, n, 64, 48, 'avc1', offs, sizes, kfs, b'\x01')
        got = hw.slice_into_video_intervals(vi, rows)
        assert got == oiv.slice_into_video_intervals(offs, sizes, kfs, n, rows)
        # properties: every row appears once, inside its interval, intervals start/end on keyframes
        flat = [r for _, rs in got for r in rs]
        assert flat == rows
        for (s, e), rs in got:
            assert s in kfs and (e in kfs or e == n) and all(s <= r < e for r in rs)


def test_video_index_serialization_is_protobuf_compatible(emu):
    """wire-compat with hwang/hwang_descriptors.proto:5-15, checked with the python protobuf runtime"""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = 'hwang_descriptors.proto'
    fd.package = 'hwang.proto'
    fd.syntax = 'proto3'
    m = fd.message_type.add()
    m.name = 'VideoIndex'
    T = descriptor_pb2.FieldDescriptorProto
    for name, num, typ, rep in (('timescale', 7, T.TYPE_UINT32, False), ('duration', 8, T.TYPE_UINT64, False),
                                ('frame_width', 1, T.TYPE_UINT32, False), ('frame_height', 2, T.TYPE_UINT32, False),
                                ('format', 9, T.TYPE_STRING, False), ('sample_offsets', 3, T.TYPE_UINT64, True),
                                ('sample_sizes', 4, T.TYPE_UINT64, True), ('keyframe_indices', 5, T.TYPE_UINT64, True),
                                ('metadata_bytes', 6, T.TYPE_BYTES, False)):
        f = m.field.add()
        f.name, f.number, f.type = name, num, typ
        f.label = T.LABEL_REPEATED if rep else T.LABEL_OPTIONAL
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    cls = message_factory.GetMessageClass(pool.FindMessageTypeByName('hwang.proto.VideoIndex'))
    offs = [0, 300, 70000, 1 << 33]
    sizes = [300, 5, 128, 9]
    vi = hw.VideoIndex.create(90000, 1 << 35, 1920, 1080, 'avc1', offs, sizes, [0, 2], bytes(range(40)))
    blob = vi.serialize()
    msg = cls()
    msg.ParseFromString(blob)
    assert list(msg.sample_offsets) == offs and list(msg.sample_sizes) == sizes and list(msg.keyframe_indices) == [0, 2]
    assert (msg.timescale, msg.duration, msg.frame_width, msg.frame_height, msg.format) == (90000, 1 << 35, 1920, 1080, 'avc1')
    assert msg.metadata_bytes == bytes(range(40))
    assert msg.SerializeToString() == blob  # byte-identical to the canonical protobuf encoding
    back = hw.VideoIndex.deserialize(msg.SerializeToString())
    assert back.sample_offsets() == offs and back.keyframe_indices() == [0, 2] and back.metadata_bytes() == bytes(range(40))
    assert back.fps() == pytest.approx(4 / ((1 << 35) / 90000))


def test_indexer_error_paths(emu, built):
    mp4 = streamgen.generate(width=64, height=48, frames=4, gop=4)
    # truncated file: EOF inside moov
    cut = mp4[:200]
    ic = run_indexer(cut)
    assert ic.is_error() and ic.error_message()
    # unsupported brand
    bad = bytearray(mp4)
    for tag in (b'isom', b'iso2', b'avc1', b'mp41'):
        i = bad.find(tag, 0, 40)
        while i >= 0:
            bad[i:i + 4] = b'qt  '
            i = bad.find(tag, 0, 40)
    ic = run_indexer(bytes(bad))
    assert ic.is_error() and 'brands' in ic.error_message()
    # different pull granularities give the same index
    a = run_indexer(mp4, 1024).get_video_index()
    b = run_indexer(mp4, 64).get_video_index()
    c = run_indexer(mp4, 1 << 20).get_video_index()
    assert a.sample_offsets() == b.sample_offsets() == c.sample_offsets()


def _boxes(buf, start, end):
    out = []
    p = start
    while p + 8 <= end:
        size, typ = struct.unpack('>I4s', buf[p:p + 8])
        if size < 8:
            break
        out.append((typ, p, size))
        p += size
    return out


def _to_co64_stz2(mp4):
    """Rewrite an unfragmented ftyp/moov/mdat file so that it uses 64-bit chunk offsets (co64) and the compact
    sample-size box (stz2, 16-bit field) when every size fits; parent box sizes and chunk offsets are fixed up."""
    buf = bytearray(mp4)
    top = _boxes(buf, 0, len(buf))
    names = [t for t, _, _ in top]
    assert names.index(b'moov') < names.index(b'mdat')
    path = [b'moov', b'trak', b'mdia', b'minf', b'stbl']
    parents = []
    lo, hi = 0, len(buf)
    for name in path:
        t, p, size = next(b for b in _boxes(buf, lo, hi) if b[0] == name)
        parents.append(p)
        lo, hi = p + 8, p + size
    stbl = {t: (p, size) for t, p, size in _boxes(buf, lo, hi)}
    # stco -> co64
    p, size = stbl[b'stco']
    n = struct.unpack('>I', buf[p + 12:p + 16])[0]
    offs = struct.unpack('>%dI' % n, buf[p + 16:p + 16 + 4 * n])
    new_stco = lambda delta: struct.pack('>I4sII', 16 + 8 * n, b'co64', 0, n) + struct.pack('>%dQ' % n, *[o + delta for o in offs])
    # stsz -> stz2 when all sizes fit in 16 bits
    q, qsize = stbl[b'stsz']
    fixed, cnt = struct.unpack('>II', buf[q + 12:q + 20])
    sizes = [fixed] * cnt if fixed else list(struct.unpack('>%dI' % cnt, buf[q + 20:q + 20 + 4 * cnt]))
    new_stsz = bytes(buf[q:q + qsize])
    if max(sizes) < 65536:
        new_stsz = struct.pack('>I4sI3xBI', 20 + 2 * cnt, b'stz2', 0, 16, cnt) + struct.pack('>%dH' % cnt, *sizes)
    delta = (16 + 8 * n - size) + (len(new_stsz) - qsize)
    pieces = sorted([(p, size, new_stco(delta)), (q, qsize, new_stsz)], reverse=True)
    for pos, old, new in pieces:
        buf[pos:pos + old] = new
    for pp in parents:
        s0 = struct.unpack('>I', buf[pp:pp + 4])[0]
        buf[pp:pp + 4] = struct.pack('>I', s0 + delta)
    return bytes(buf), delta


@pytest.mark.parametrize('name', NAMES)
def test_index_of_co64_stz2_variant(emu, name):
    """64-bit chunk offsets and compact sample sizes.  The rewritten file must index to the same
    table shifted by the growth of the moov box; checked against the reference's own indexer when it is built here."""
    mp4, g = load(name)
    ri = g['reference_index']
    top = [t for t, _, _ in _boxes(mp4, 0, len(mp4))]
    if b'moof' in top or b'mdat' not in top or top.index(b'moov') > top.index(b'mdat'):
        pytest.skip('fragmented or mdat-first layout')
    v64, delta = _to_co64_stz2(mp4)
    ic = run_indexer(v64)
    assert not ic.is_error(), ic.error_message()
    vi = ic.get_video_index()
    assert vi.sample_offsets() == [o + delta for o in ri['offsets']]
    assert vi.sample_sizes() == ri['sizes']
    assert vi.keyframe_indices() == ri['keyframes']
    tool = os.path.join(os.path.dirname(GOLDEN), '..', 'oracle', '_ref', 'ref_tool')
    if os.path.exists(tool):
        import subprocess, tempfile
        with tempfile.NamedTemporaryFile(suffix='.mp4') as f:
            f.write(v64); f.flush()
            ref = json.loads(subprocess.run([tool, 'index', f.name], capture_output=True, text=True, timeout=60).stdout)
        assert 'error' not in ref, ref
        assert (ref['offsets'], ref['sizes'], ref['keyframes']) == (vi.sample_offsets(), vi.sample_sizes(), vi.keyframe_indices())


def _mdat_first(mp4):
    """ftyp/moov/mdat -> ftyp/mdat/moov (the layout most encoders write without "faststart"): chunk offsets move
    back by the size of the moov box."""
    top = _boxes(mp4, 0, len(mp4))
    by = {t: (p, size) for t, p, size in top}
    assert [t for t, _, _ in top] == [b'ftyp', b'moov', b'mdat']
    mp, ms = by[b'moov']
    moov = bytearray(mp4[mp:mp + ms])
    # find stco inside moov and shift its entries
    lo, hi = 8, ms
    for name in (b'trak', b'mdia', b'minf', b'stbl'):
        t, p, size = next(b for b in _boxes(moov, lo, hi) if b[0] == name)
        lo, hi = p + 8, p + size
    t, p, size = next(b for b in _boxes(moov, lo, hi) if b[0] == b'stco')
    n = struct.unpack('>I', moov[p + 12:p + 16])[0]
    offs = struct.unpack('>%dI' % n, moov[p + 16:p + 16 + 4 * n])
    moov[p + 16:p + 16 + 4 * n] = struct.pack('>%dI' % n, *[o - ms for o in offs])
    fp, fs = by[b'ftyp']
    dp, ds = by[b'mdat']
    return mp4[fp:fp + fs] + mp4[dp:dp + ds] + bytes(moov), -ms


@pytest.mark.parametrize('name', NAMES)
def test_index_of_mdat_first_variant(emu, name):
    """moov after mdat: the pull parser must skip the media data without asking for it, then read the moov box."""
    mp4, g = load(name)
    ri = g['reference_index']
    if [t for t, _, _ in _boxes(mp4, 0, len(mp4))] != [b'ftyp', b'moov', b'mdat']:
        pytest.skip('not an ftyp/moov/mdat file')
    v, delta = _mdat_first(mp4)
    ic = hw.MP4IndexCreator(len(v))
    off, size, asked = 0, 1024, 0
    while not ic.is_done():
        asked += size
        _, off, size = ic.feed(v[off:off + size], size)
    assert not ic.is_error(), ic.error_message()
    vi = ic.get_video_index()
    assert vi.sample_offsets() == [o + delta for o in ri['offsets']]
    assert vi.sample_sizes() == ri['sizes'] and vi.keyframe_indices() == ri['keyframes']
    tool = os.path.join(os.path.dirname(GOLDEN), '..', 'oracle', '_ref', 'ref_tool')
    if os.path.exists(tool):
        import subprocess, tempfile
        with tempfile.NamedTemporaryFile(suffix='.mp4') as f:
            f.write(v); f.flush()
            ref = json.loads(subprocess.run([tool, 'index', f.name], capture_output=True, text=True, timeout=60).stdout)
        assert 'error' not in ref, ref
        assert (ref['offsets'], ref['sizes'], ref['keyframes']) == (vi.sample_offsets(), vi.sample_sizes(), vi.keyframe_indices())


def _box(typ, payload):
    return struct.pack('>I4s', 8 + len(payload), typ) + payload


def _audio_trak(track_id=1, samples=5):
    """A minimal sound track (handler 'soun', one chunk of `samples` 100-byte samples at offset 0x1000)."""
    tkhd = _box(b'tkhd', struct.pack('>IIIII', 7, 0, 0, track_id, 0) + bytes(60))
    mdhd = _box(b'mdhd', struct.pack('>IIIIIHH', 0, 0, 0, 44100, 44100 * 2, 0x55C4, 0))
    hdlr = _box(b'hdlr', struct.pack('>II4s', 0, 0, b'soun') + bytes(12) + b'snd\0')
    mp4a = _box(b'mp4a', bytes(6) + struct.pack('>H', 1) + bytes(8) + struct.pack('>HHHHI', 2, 16, 0, 0, 44100 << 16))
    stbl = _box(b'stbl', _box(b'stsd', struct.pack('>II', 0, 1) + mp4a) + _box(b'stts', struct.pack('>IIII', 0, 1, samples, 1024)) +
                _box(b'stsc', struct.pack('>IIIII', 0, 1, 1, samples, 1)) + _box(b'stsz', struct.pack('>III', 0, 100, samples)) +
                _box(b'stco', struct.pack('>III', 0, 1, 0x1000)))
    minf = _box(b'minf', _box(b'smhd', bytes(8)) + _box(b'dinf', _box(b'dref', struct.pack('>II', 0, 1) + _box(b'url ', struct.pack('>I', 1)))) + stbl)
    return _box(b'trak', tkhd + _box(b'mdia', mdhd + hdlr + minf))


@pytest.mark.parametrize('name', NAMES)
def test_index_of_a_file_with_a_sound_track_first_extra_boxes_and_a_64_bit_mdat(emu, name):
    """What muxers other than ours write around the video track: a sound track AHEAD of it inside moov, `free` / `udta`
    boxes, an `edts` box inside the video track, and an mdat box with a 64-bit size field.  The video index must be the
    same table, shifted; checked against the reference's own indexer when it is built here."""
    mp4, g = load(name)
    ri = g['reference_index']
    buf = bytes(mp4)
    top = _boxes(buf, 0, len(buf))
    names = [t for t, _, _ in top]
    if b'moof' in names or b'mdat' not in names or names.index(b'moov') > names.index(b'mdat'):
        pytest.skip('fragmented or mdat-first layout')
    (_, mp, msize), (_, dp, dsize) = next(b for b in top if b[0] == b'moov'), next(b for b in top if b[0] == b'mdat')
    moov_kids = _boxes(buf, mp + 8, mp + msize)
    _, tp, tsize = next(b for b in moov_kids if b[0] == b'trak')
    edts = _box(b'edts', _box(b'elst', struct.pack('>IIIiI', 0, 1, 1000, 0, 0x10000)))
    trak_kids = _boxes(buf, tp + 8, tp + tsize)
    _, kp, ksize = trak_kids[0]  # tkhd first, edts right behind it
    video_trak = _box(b'trak', buf[tp + 8:kp + ksize] + edts + buf[kp + ksize:tp + tsize])
    new_moov_payload = buf[mp + 8:tp] + _audio_trak(track_id=7) + video_trak + buf[tp + tsize:mp + msize] + \
        _box(b'udta', _box(b'meta', bytes(4) + _box(b'hdlr', bytes(24))))
    free = _box(b'free', bytes(37))
    new_mdat = struct.pack('>I4sQ', 1, b'mdat', dsize + 8) + buf[dp + 8:dp + dsize]  # largesize form
    head = buf[:mp] + _box(b'moov', new_moov_payload) + buf[mp + msize:dp] + free
    delta = len(head) + 16 - (dp + 8)  # where the media data starts now minus where it started
    out = bytearray(head + new_mdat + buf[dp + dsize:])
    # the video track's chunk offsets move by delta (they sit in the last stco of the file: the sound track's comes first)
    sp = bytes(out).rfind(b'stco')
    n = struct.unpack('>I', out[sp + 8:sp + 12])[0]
    offs = struct.unpack('>%dI' % n, out[sp + 12:sp + 12 + 4 * n])
    out[sp + 12:sp + 12 + 4 * n] = struct.pack('>%dI' % n, *[o + delta for o in offs])
    out = bytes(out)
    ic = run_indexer(out)
    assert not ic.is_error(), ic.error_message()
    vi = ic.get_video_index()
    assert vi.sample_offsets() == [o + delta for o in ri['offsets']]
    assert vi.sample_sizes() == ri['sizes'] and vi.keyframe_indices() == ri['keyframes']
    assert (vi.frame_width(), vi.frame_height()) == (ri['width'], ri['height'])
    for o, s in zip(vi.sample_offsets(), vi.sample_sizes()):  # the table points at the same bytes as before
        assert out[o:o + s] == buf[o - delta:o - delta + s]
    tool = os.path.join(os.path.dirname(GOLDEN), '..', 'oracle', '_ref', 'ref_tool')
    if os.path.exists(tool):
        import subprocess, tempfile
        with tempfile.NamedTemporaryFile(suffix='.mp4') as f:
            f.write(out); f.flush()
            ref = json.loads(subprocess.run([tool, 'index', f.name], capture_output=True, text=True, timeout=60).stdout)
        if 'error' not in ref:  # the reference indexer may refuse what it does not know; when it answers, the answers must agree
            assert (ref['offsets'], ref['sizes'], ref['keyframes']) == (vi.sample_offsets(), vi.sample_sizes(), vi.keyframe_indices())


@pytest.mark.parametrize('name', NAMES)
def test_index_of_an_hevc_sample_entry(emu, name):
    """The reference indexes HEVC files too (hev1 / hvcC, mp4_index_creator.cpp:454-469; its decode tests use one): the
    index -- table, format string, configuration record -- does not depend on the codec.  Decoding such a file is
    refused with the reference's "Unsupported video codec" wording."""
    mp4, g = load(name)
    ri = g['reference_index']
    assert mp4.count(b'avc1') >= 1 and mp4.count(b'avcC') == 1
    hevc = mp4.replace(b'avcC', b'hvcC')
    i = hevc.find(b'stsd')
    j = hevc.find(b'avc1', i)
    hevc = hevc[:j] + b'hev1' + hevc[j + 4:]
    ic = run_indexer(hevc)
    assert not ic.is_error(), ic.error_message()
    vi = ic.get_video_index()
    assert vi.format() == 'hev1'
    assert vi.sample_offsets() == ri['offsets'] and vi.sample_sizes() == ri['sizes'] and vi.keyframe_indices() == ri['keyframes']
    assert vi.metadata_bytes() == hw.index_video(io.BytesIO(mp4)).metadata_bytes()
    tool = os.path.join(os.path.dirname(GOLDEN), '..', 'oracle', '_ref', 'ref_tool')
    if os.path.exists(tool):
        import subprocess, tempfile
        with tempfile.NamedTemporaryFile(suffix='.mp4') as f:
            f.write(hevc); f.flush()
            ref = json.loads(subprocess.run([tool, 'index', f.name], capture_output=True, text=True, timeout=60).stdout)
        assert 'error' not in ref, ref
        assert (ref['offsets'], ref['sizes'], ref['keyframes'], ref['format']) == (vi.sample_offsets(), vi.sample_sizes(), vi.keyframe_indices(), 'hev1')
    dec = hw.VideoDecoder(0)
    with pytest.raises(RuntimeError, match='Unsupported video codec'):
        dec.configure(vi.frame_width(), vi.frame_height(), vi.format(), vi.metadata_bytes())
