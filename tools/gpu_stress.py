"""Stress: many automaton passes over a mid-size clip with varying chunk sizes and batch sizes, checking every frame's
checksum against the first pass (and the first pass against libavcodec).  Prints any decoder error message."""
import hashlib, io, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import hwang_b200 as hw
from hwang_b200 import _lib
_lib.use_library(_lib.PRODUCT_LIB)
import hwb_testutil as util
from oracle import ffmpeg_oracle as fo
kw = dict(width=640, height=480, frames=360, gop=12, profile=1, bframes=1, num_ref=2, seed=99, qp=30)
mp4, index, samples, kf = util.make_clip(**kw)
ref = [hashlib.md5(fo.yuv420_to_rgb24(*f).tobytes()).hexdigest() for f in util.oracle_frames(index, samples, kf)]
offs, sizes = index.sample_offsets(), index.sample_sizes()
n = len(samples)
L = _lib.lib()
rng = random.Random(1)
fails = 0
t0 = time.time()
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    chunk = rng.choice([1, 12, 30, 60, 100, 200, 4096])
    batch = rng.choice([1, 7, 8, 64, 360])
    os.environ['HWB_CHUNK_PICTURES'] = str(chunk)
    auto = hw.DecoderAutomata(hw.DeviceHandle(hw.DeviceType.GPU, 0), 1, hw.VideoDecoderType.B200)
    ed = hw.EncodedData()
    ed.width, ed.height, ed.format = kw['width'], kw['height'], index.format()
    ed.start_keyframe, ed.end_keyframe = 0, n
    ed.sample_offsets = [o - offs[0] for o in offs]; ed.sample_sizes = sizes; ed.keyframes = sorted(index.keyframe_indices()); ed.valid_frames = list(range(n))
    ed.encoded_video = mp4[offs[0]:offs[-1] + sizes[-1]]
    for rep in range(2):
        try:
            auto.initialize([ed], index.metadata_bytes())
            got = []
            while len(got) < n:
                k = min(batch, n - len(got))
                got += [hashlib.md5(np.asarray(f).tobytes()).hexdigest() for f in auto.get_frames(index, k)]
            bad = [i for i in range(n) if got[i] != ref[i]]
            if bad:
                fails += 1
                print('MISMATCH it', it, 'chunk', chunk, 'batch', batch, 'rep', rep, 'frames', bad[:8], len(bad), flush=True)
        except Exception as e:
            fails += 1
            print('ERROR it', it, 'chunk', chunk, 'batch', batch, 'rep', rep, repr(e)[:300], flush=True)
    del auto
print('stress done: %d failures in %.1fs' % (fails, time.time() - t0))
