"""Ad-hoc GPU bisect helper: decode clips with the product library and report where they differ from libavcodec."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import hwang_b200 as hw
from hwang_b200 import _lib
_lib.use_library(os.environ.get('HWB_LIB', _lib.PRODUCT_LIB))
import hwb_testutil as util

base = dict(frames=20, gop=10, width=320, height=240, seed=11, num_ref=2, slices=3, qp_jitter=3, ipcm_per_100k=2000, intra_in_p_pct=10)
variants = {
    'full': {}, 'slices1': dict(slices=1), 'noipcm': dict(ipcm_per_100k=0), 'nointra': dict(intra_in_p_pct=0),
    'noipcm_slices1': dict(ipcm_per_100k=0, slices=1), 'ionly': dict(gop=1, frames=4), 'ionly_noipcm': dict(gop=1, frames=4, ipcm_per_100k=0),
    'nodeblock': dict(deblock=1), 'cabac': dict(profile=1),
}
for name, v in variants.items():
    kw = dict(base); kw.update(v)
    mp4, index, samples, kf = util.make_clip(**kw)
    ref = util.oracle_frames(index, samples, kf)
    for rep in range(2):
        got, dec = util.decode_yuv(index, samples, kf)
        W, H = kw['width'], kw['height']
        bad = []
        for i, (g, r) in enumerate(zip(got, ref)):
            e = util.flat(r)
            if not np.array_equal(g, e):
                d = np.nonzero(g != e)[0]
                first = int(d[0])
                where = ('Y', first % W, first // W) if first < W * H else ('C', first - W * H, 0)
                bad.append((i, len(d), where, (where[1] // 16, where[2] // 16)))
        print(name, 'rep', rep, 'bad frames:', len(bad), bad[:4], flush=True)
