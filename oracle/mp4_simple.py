"""ORACLE helper (test infrastructure only): minimal ISO-BMFF walker, a Python restatement of what
the reference's MP4IndexCreator extracts (hwang/mp4_index_creator.cpp:265-453, 641-711): sample
offsets/sizes, sync samples, avcC, width/height.  Handles the unfragmented stsz/stsc/stco|co64/stss
layout and moof/traf/trun fragments."""
import struct


def _boxes(buf, start, end):
    off = start
    while off + 8 <= end:
        size, typ = struct.unpack('>I4s', buf[off:off + 8])
        hdr = 8
        if size == 1:
            size = struct.unpack('>Q', buf[off + 8:off + 16])[0]
            hdr = 16
        elif size == 0:
            size = end - off
        yield typ, off + hdr, off + size, off
        off += size


def _find(buf, start, end, path):
    for typ, s, e, _ in _boxes(buf, start, end):
        if typ == path[0]:
            if len(path) == 1:
                return s, e
            r = _find(buf, s, e, path[1:])
            if r:
                return r
    return None


def index_mp4(buf):
    buf = bytes(buf)
    out = {'offsets': [], 'sizes': [], 'keyframes': []}
    moov = _find(buf, 0, len(buf), [b'moov'])
    stbl = _find(buf, moov[0], moov[1], [b'trak', b'mdia', b'minf', b'stbl'])
    mdhd = _find(buf, moov[0], moov[1], [b'trak', b'mdia', b'mdhd'])
    out['timescale'], out['duration'] = struct.unpack('>II', buf[mdhd[0] + 12:mdhd[0] + 20])
    stsd = _find(buf, stbl[0], stbl[1], [b'stsd'])
    e0 = stsd[0] + 8  # first sample entry
    esz, fmt = struct.unpack('>I4s', buf[e0:e0 + 8])
    out['format'] = fmt.decode()
    out['width'], out['height'] = struct.unpack('>HH', buf[e0 + 32:e0 + 36])
    for typ, s, e, _ in _boxes(buf, e0 + 86, e0 + esz):
        if typ == b'avcC':
            out['avcc'] = buf[s:e]
    stsz = _find(buf, stbl[0], stbl[1], [b'stsz'])
    fixed, n = struct.unpack('>II', buf[stsz[0] + 4:stsz[0] + 12])
    sizes = [fixed] * n if fixed else list(struct.unpack('>%dI' % n, buf[stsz[0] + 12:stsz[0] + 12 + 4 * n]))
    if n:
        stsc = _find(buf, stbl[0], stbl[1], [b'stsc'])
        ne = struct.unpack('>I', buf[stsc[0] + 4:stsc[0] + 8])[0]
        runs = [struct.unpack('>III', buf[stsc[0] + 8 + 12 * i:stsc[0] + 20 + 12 * i]) for i in range(ne)]
        co = _find(buf, stbl[0], stbl[1], [b'stco'])
        if co:
            nc = struct.unpack('>I', buf[co[0] + 4:co[0] + 8])[0]
            chunks = list(struct.unpack('>%dI' % nc, buf[co[0] + 8:co[0] + 8 + 4 * nc]))
        else:
            co = _find(buf, stbl[0], stbl[1], [b'co64'])
            nc = struct.unpack('>I', buf[co[0] + 4:co[0] + 8])[0]
            chunks = list(struct.unpack('>%dQ' % nc, buf[co[0] + 8:co[0] + 8 + 8 * nc]))
        si = 0
        for ci in range(nc):
            spc = 0
            for (first, per, _d) in runs:
                if ci + 1 >= first:
                    spc = per
            off = chunks[ci]
            for _ in range(spc):
                if si >= n:
                    break
                out['offsets'].append(off)
                out['sizes'].append(sizes[si])
                off += sizes[si]
                si += 1
        stss = _find(buf, stbl[0], stbl[1], [b'stss'])
        if stss:
            nk = struct.unpack('>I', buf[stss[0] + 4:stss[0] + 8])[0]
            out['keyframes'] = [k - 1 for k in struct.unpack('>%dI' % nk, buf[stss[0] + 8:stss[0] + 8 + 4 * nk])]
        else:
            out['keyframes'] = list(range(n))
    # fragments
    for typ, s, e, box_off in _boxes(buf, 0, len(buf)):
        if typ != b'moof':
            continue
        traf = _find(buf, s, e, [b'traf'])
        base = box_off
        def_size = 0
        for t2, s2, e2, _ in _boxes(buf, traf[0], traf[1]):
            if t2 == b'tfhd':
                flags = struct.unpack('>I', buf[s2:s2 + 4])[0] & 0xFFFFFF
                p = s2 + 8
                if flags & 1:
                    base = struct.unpack('>Q', buf[p:p + 8])[0]; p += 8
                if flags & 2:
                    p += 4
                if flags & 8:
                    p += 4
                if flags & 0x10:
                    def_size = struct.unpack('>I', buf[p:p + 4])[0]; p += 4
            elif t2 == b'trun':
                flags = struct.unpack('>I', buf[s2:s2 + 4])[0] & 0xFFFFFF
                cnt = struct.unpack('>I', buf[s2 + 4:s2 + 8])[0]
                p = s2 + 8
                off = base
                if flags & 1:
                    off = base + struct.unpack('>i', buf[p:p + 4])[0]; p += 4
                first_flags = None
                if flags & 4:
                    first_flags = struct.unpack('>I', buf[p:p + 4])[0]; p += 4
                for i in range(cnt):
                    sz = def_size
                    sflags = None
                    if flags & 0x100:
                        p += 4
                    if flags & 0x200:
                        sz = struct.unpack('>I', buf[p:p + 4])[0]; p += 4
                    if flags & 0x400:
                        sflags = struct.unpack('>I', buf[p:p + 4])[0]; p += 4
                    if flags & 0x800:
                        p += 4
                    if i == 0 and first_flags is not None:
                        sflags = first_flags
                    if sflags is not None and not (sflags & 0x10000):
                        out['keyframes'].append(len(out['sizes']))
                    out['offsets'].append(off)
                    out['sizes'].append(sz)
                    off += sz
    return out
