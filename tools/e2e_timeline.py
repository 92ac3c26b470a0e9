"""When does each batch of frames reach the host?  One dense pass of the benchmark clip through DecoderAutomata.get_frames
(64 frames per call, page-locked destination), printing the wall-clock time of every call's return: pipeline fill
(first batch), steady-state rate, tail.  Usage: python tools/e2e_timeline.py [frames] [reps]"""
import io, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import hwang_b200 as hw
from hwang_b200 import _lib
_lib.use_library(_lib.PRODUCT_LIB)
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mp4 = bench.get_clip(frames)
index = hw.index_video(io.BytesIO(mp4))
intervals = hw.api.encoded_intervals(io.BytesIO(mp4), index, list(range(frames)))
L = _lib.lib()
fs = bench.W * bench.H * 3
batch = 64
pinned = hw.api.PinnedBuffer(fs * batch)
auto = hw.DecoderAutomata(hw.DeviceHandle(hw.DeviceType.GPU, 0), 1, hw.VideoDecoderType.B200)
for rep in range(reps):
    t0 = time.perf_counter()
    auto.initialize(intervals, index.metadata_bytes())
    t_init = time.perf_counter() - t0
    stamps, done = [], 0
    while done < frames:
        k = min(batch, frames - done)
        assert L.hwb_automata_get_frames(auto._h, pinned.ptr, k) == 0
        done += k
        stamps.append((done, round((time.perf_counter() - t0) * 1000, 1)))
    total = stamps[-1][1]
    # steady-state rate between 25% and 90% of the frames
    a = next(s for s in stamps if s[0] >= frames // 4)
    b = next(s for s in stamps if s[0] >= frames * 9 // 10)
    print(json.dumps({'rep': rep, 'init_ms': round(t_init * 1000, 1), 'first_batch_ms': stamps[0][1], 'total_ms': total, 'fps': round(frames / total * 1000),
                      'steady_fps': round((b[0] - a[0]) / (b[1] - a[1]) * 1000), 'every_8th_batch': stamps[::8],
                      'env': {k: v for k, v in os.environ.items() if k.startswith('HWB_')}}), flush=True)
