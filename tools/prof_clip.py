"""Profiling driver for one synthetic clip shape (run under ncu): python tools/prof_clip.py <frames> <gop> [reps]"""
import io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import hwang_b200 as hw
from hwang_b200 import _lib, build
from hwang_b200.testing import streamgen
_lib.use_library(_lib.PRODUCT_LIB)
build.build_gen()
frames, gop = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
mp4 = streamgen.generate(frames=frames, **{**bench.CLIP_KW, 'gop': gop})
index = hw.index_video(io.BytesIO(mp4))
offs, sizes, kf = index.sample_offsets(), index.sample_sizes(), set(index.keyframe_indices())
samples = [mp4[o:o + s] for o, s in zip(offs, sizes)]
dec = hw.VideoDecoder(0)
dec.set_chunk_pictures(1 << 30)
for r in range(reps):
    dec.configure(bench.W, bench.H, index.format(), index.metadata_bytes())
    for i, s in enumerate(samples):
        dec.feed(s, i in kf)
    dec.feed(None); dec.flush()
    n = 0
    while n < frames:
        if dec.frames_ready() != 0:
            dec.get_frame_device(); n += 1
        else:
            time.sleep(0.001)
    dec.wait_until_frames_copied()
print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in dec.stats().items()})
