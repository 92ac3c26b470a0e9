"""Tiny decode through the product library (run under compute-sanitizer before spending GPU time on anything else)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from hwang_b200 import _lib
_lib.use_library(_lib.PRODUCT_LIB)
import hwb_testutil as util
for kw in (dict(frames=6, gop=3, width=64, height=48, profile=1, bframes=1, seed=5, slices=2),
           dict(frames=4, gop=4, width=64, height=48, profile=0, seed=6, ipcm_per_100k=5000),
           dict(frames=5, gop=5, width=64, height=48, profile=2, bframes=2, seed=7, direct_spatial=0),
           dict(frames=5, gop=5, width=64, height=48, profile=1, seed=8),    # Main, I/P only: entropy_cabac_ip4_kernel
           dict(frames=5, gop=5, width=64, height=48, profile=2, seed=9)):   # High, I/P only: entropy_cabac_ip_kernel
    util.assert_yuv_parity(kw)
    print('ok', kw, flush=True)
