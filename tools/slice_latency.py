"""Per-slice entropy latency probe: decode clips whose pictures are all of one kind, so that the entropy kernel's
duration is one slice's latency (every slice of the chunk runs concurrently, one warp each).
  python tools/slice_latency.py [frames]
"""
import io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import hwang_b200 as hw
from hwang_b200 import _lib, build
from hwang_b200.testing import streamgen
_lib.use_library(_lib.PRODUCT_LIB)
build.build_gen()


def run(name, **kw):
    mp4 = streamgen.generate(**kw)
    index = hw.index_video(io.BytesIO(mp4))
    offs, sizes, kf = index.sample_offsets(), index.sample_sizes(), set(index.keyframe_indices())
    samples = [mp4[o:o + s] for o, s in zip(offs, sizes)]
    dec = hw.VideoDecoder(0)
    dec.set_chunk_pictures(1 << 30)
    best = None
    for r in range(3):
        s0 = dec.stats()
        dec.configure(kw['width'], kw['height'], index.format(), index.metadata_bytes())
        for i, s in enumerate(samples):
            dec.feed(s, i in kf)
        dec.feed(None); dec.flush()
        n = 0
        while n < len(samples):
            if dec.frames_ready() != 0:
                dec.get_frame_device(); n += 1
            else:
                time.sleep(0.001)
        dec.wait_until_frames_copied()
        s1 = dec.stats()
        d = {k: s1[k] - s0[k] for k in ('entropy_ms', 'picture_ms', 'wall_ms')}
        if best is None or d['entropy_ms'] < best['entropy_ms']:
            best = d
    print('%-28s frames %4d  avg bytes/frame %7d  entropy %8.2f ms  picture kernel %7.2f ms' % (
        name, len(samples), sum(sizes) // len(sizes), best['entropy_ms'], best['picture_ms']), flush=True)


base = dict(bench.CLIP_KW)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
run('all-I (gop 1)', frames=n, **{**base, 'gop': 1})
run('1 I + P (gop = frames)', frames=n, **{**base, 'gop': n})
run('bench shape gop 30', frames=max(n, 60), **base)
