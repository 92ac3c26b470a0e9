"""Sparse-retrieval timing (the "sparse ms/frame" half of BASELINE.json's metric), through hwang.Decoder.retrieve.
  config 3: 1080p High (8x8 transform, B pictures), every 17th frame
  config 4: 3840x2160 High, GOP 250, 64 seeded random rows
Prints one JSON line per config with ms per returned frame for the CUDA path and for the reference CPU path
(libavcodec threads=1 + sws_scale, decoding each interval from its keyframe and dropping unwanted frames as the
reference does, decoder_automata.cpp:235)."""
import io, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hwang_b200 as hw
from hwang_b200 import _lib
from hwang_b200.testing import streamgen
from oracle import ffmpeg_oracle as fo
_lib.use_library(_lib.PRODUCT_LIB)

def clip(name, **kw):
    p = os.path.join(ROOT, 'tests', '_cache', name + '.mp4')
    os.makedirs(os.path.dirname(p), exist_ok=True)
    if not os.path.exists(p):
        t = time.time()
        open(p, 'wb').write(streamgen.generate(**kw))
        sys.stderr.write('generated %s in %.1fs\n' % (name, time.time() - t))
    return open(p, 'rb').read()

def cpu_sparse(mp4, index, rows):
    ivs = hw.slice_into_video_intervals(index, rows)
    offs, sizes, kfs = index.sample_offsets(), index.sample_sizes(), set(index.keyframe_indices())
    nls, sps, pps = fo.parse_avcc(index.metadata_bytes())
    sws = fo.SwsRgb24(index.frame_width(), index.frame_height())
    t = time.perf_counter()
    n = 0
    for (s, e), valid in ivs:
        want = set(valid)
        dec = fo.FFmpegH264(threads=1)
        k = [s]
        def sink(yuv):
            if k[0] in want:
                sws(*yuv)
            k[0] += 1
        for i in range(s, e):
            dec.send(fo.avcc_to_annexb(mp4[offs[i]:offs[i] + sizes[i]], nls, sps, pps, i in kfs), sink)
            if k[0] > valid[-1]:
                break  # the reference feeder stops once the consumer has its frames (decoder_automata.cpp:287)
        else:
            dec.send(None, sink)
        dec.close()
        n += len(valid)
    return (time.perf_counter() - t) * 1000.0 / n

def run(name, rows, **kw):
    mp4 = clip(name, **kw)
    index = hw.index_video(io.BytesIO(mp4))
    dec = hw.Decoder(io.BytesIO(mp4), video_index=index)
    dec.retrieve(rows[:2])  # warm-up
    best = 1e9
    for _ in range(3):
        t = time.perf_counter()
        frames = dec.retrieve(rows)
        best = min(best, (time.perf_counter() - t) * 1000.0 / len(rows))
    cpu = cpu_sparse(mp4, index, rows)
    print(json.dumps({'config': name, 'rows': len(rows), 'frames_in_clip': index.frames(), 'intervals': len(hw.slice_into_video_intervals(index, rows)),
                      'gpu_ms_per_returned_frame': round(best, 3), 'cpu_ms_per_returned_frame_1thread': round(cpu, 3),
                      'bits_per_frame': round(8 * len(mp4) / index.frames())}), flush=True)

if __name__ == '__main__':
    n3 = int(os.environ.get('HWB_SPARSE3_FRAMES', '600'))
    run('sparse3_1080p_high_b_%d' % n3, list(range(0, n3, 17)), width=1920, height=1080, frames=n3, gop=30, profile=2, bframes=2, num_ref=3, weighted=2, seed=3, qp=27)
    n4 = int(os.environ.get('HWB_SPARSE4_FRAMES', '500'))
    rng = np.random.default_rng(0)
    rows = sorted(set(int(x) for x in rng.integers(0, n4, 64)))
    run('sparse4_4k_high_gop250_%d' % n4, rows, width=3840, height=2160, frames=n4, gop=250, profile=2, bframes=1, num_ref=2, seed=4, qp=30)
