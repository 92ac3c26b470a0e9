"""Tiny decode through the product library (run under compute-sanitizer before spending GPU time on anything else)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from hwang_b200 import _lib
_lib.use_library(_lib.PRODUCT_LIB)
import hwb_testutil as util
ONLY = os.environ.get('HWB_SANITY_ONLY', '')  # '', 'yuv', 'rgb', 'rgb0' (first RGB clip only), 'rgb0yuv' (that clip through the planar-output path)
RGB_CLIPS = (dict(frames=6, gop=3, width=64, height=48, profile=1, seed=10), dict(frames=6, gop=3, width=72, height=48, profile=2, bframes=1, seed=11))
if ONLY == 'rgb0yuv':
    util.assert_yuv_parity(RGB_CLIPS[0]); print('ok', RGB_CLIPS[0], flush=True)
for kw in () if ONLY not in ('', 'yuv') else (dict(frames=6, gop=3, width=64, height=48, profile=1, bframes=1, seed=5, slices=2),
           dict(frames=4, gop=4, width=64, height=48, profile=0, seed=6, ipcm_per_100k=5000),
           dict(frames=5, gop=5, width=64, height=48, profile=2, bframes=2, seed=7, direct_spatial=0),
           dict(frames=5, gop=5, width=64, height=48, profile=1, seed=8),    # Main, I/P only: entropy_cabac_ip4_kernel
           dict(frames=5, gop=5, width=64, height=48, profile=2, seed=9)):   # High, I/P only: entropy_cabac_ip_kernel
    util.assert_yuv_parity(kw)
    print('ok', kw, flush=True)

# RGB24 through the picture kernel's fused writeback: a width whose rows are 16-byte aligned (vector path) and one whose
# rows are not (byte path), every frame against the oracle's swscale arithmetic
import io
import numpy as np
import hwang_b200 as hw
from oracle import ffmpeg_oracle as fo
for kw in () if ONLY not in ('', 'rgb', 'rgb0') else RGB_CLIPS[:1 if ONLY == 'rgb0' else 2]:
    mp4, index, samples, kf = util.make_clip(**kw)
    ref = util.oracle_frames(index, samples, kf)
    frames = hw.Decoder(io.BytesIO(mp4), video_index=index).retrieve(list(range(kw['frames'])))
    for r, f in enumerate(frames):
        assert np.array_equal(np.asarray(f), fo.yuv420_to_rgb24(*ref[r])), 'rgb frame %d of %s' % (r, kw)
    print('ok rgb', kw, flush=True)
